// tchem::polynomial::PolynomialSet - drop-in for include/tchem/polynomial/PolynomialSet.hpp of Torch-Chemistry:
// construction (explicit terms, or all terms up to an order), value, Jacobian and second-order Jacobian,
// evaluated by the CUDA library on the SAPSet kernels (a polynomial set is a SAPSet with one irreducible).
// Batched: x may be [dimension] or [B][dimension]. rotation / translation are set-up time utilities outside
// the B200 hot path and are not provided.
#ifndef tchem_polynomial_PolynomialSet_hpp
#define tchem_polynomial_PolynomialSet_hpp

#include <memory>

#include <tchem/polynomial/Polynomial.hpp>

namespace tchem { namespace polynomial {

class SAPSet;

class PolynomialSet {
    private:
        std::vector<Polynomial> polynomials_;
        size_t dimension_ = 0;
        size_t max_order_ = 0;
        std::vector<std::vector<const Polynomial *>> orders_;
        void construct_orders_();
        // the same monomials as a one-irreducible SAPSet (owns the device tables)
        mutable std::shared_ptr<SAPSet> engine_;
        const SAPSet & engine() const;
        void check_(const at::Tensor & x, const char * who) const;
    public:
        PolynomialSet();
        PolynomialSet(const std::vector<Polynomial> & _polynomials, const size_t & _dimension);
        PolynomialSet(const size_t & _dimension, const size_t & _order);
        ~PolynomialSet();

        const std::vector<Polynomial> & polynomials() const;
        const size_t & dimension() const;
        const size_t & max_order() const;
        const std::vector<std::vector<const Polynomial *>> & orders() const;
        const Polynomial & operator[](const size_t & index) const;

        std::vector<at::Tensor> views(const at::Tensor & x) const;

        at::Tensor operator()(const at::Tensor & x) const;
        at::Tensor Jacobian(const at::Tensor & x) const;
        at::Tensor Jacobian_(const at::Tensor & x) const;
        at::Tensor Jacobian2nd(const at::Tensor & x) const;
        at::Tensor Jacobian2nd_(const at::Tensor & x) const;
};

} }

#endif
