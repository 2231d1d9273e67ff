// tchem::polynomial::Polynomial - drop-in for include/tchem/polynomial/Polynomial.hpp of Torch-Chemistry:
// one monomial P(x) = x[coords_[0]] * x[coords_[1]] * ... Its evaluators run through a one-term
// PolynomialSet on the CUDA library (no CPU arithmetic path).
#ifndef tchem_polynomial_Polynomial_hpp
#define tchem_polynomial_Polynomial_hpp

#include <torch/torch.h>

namespace tchem { namespace polynomial {

class Polynomial {
    private:
        std::vector<size_t> coords_;                                   // sorted descending
        std::vector<std::pair<size_t, size_t>> uniques_orders_;        // (coordinate, power), ascending coordinate
    public:
        Polynomial();
        Polynomial(const std::vector<size_t> & _coords);
        Polynomial(const std::vector<std::pair<size_t, size_t>> _uniques_orders);
        ~Polynomial();

        const std::vector<size_t> & coords() const;
        const std::vector<std::pair<size_t, size_t>> & uniques_orders() const;
        size_t order() const;

        at::Tensor operator()(const at::Tensor & x) const;
        at::Tensor gradient(const at::Tensor & x) const;
        at::Tensor gradient_(const at::Tensor & x) const;
        at::Tensor Hessian(const at::Tensor & x) const;
        at::Tensor Hessian_(const at::Tensor & x) const;
};

} }

#endif
