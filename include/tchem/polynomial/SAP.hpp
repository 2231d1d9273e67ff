// tchem::polynomial::SAP - drop-in for include/tchem/polynomial/SAP.hpp of Torch-Chemistry,
// value, gradient and Hessian of one monomial (per-term helpers in at:: tensor arithmetic, like the
// reference's autograd-friendly variants; the batched work goes through SAPSet).
#ifndef tchem_polynomial_SAP_hpp
#define tchem_polynomial_SAP_hpp

#include <torch/torch.h>

#include <CppLibrary/utility.hpp>

namespace tchem { namespace polynomial {

// SAP(x) = product over coords_ of x[irreducible][index]
class SAP {
    private:
        std::vector<std::pair<size_t, size_t>> coords_;                              // sorted descending
        std::vector<std::pair<std::pair<size_t, size_t>, size_t>> uniques_orders_;   // (coordinate, power)
        void build_uniques_();
    public:
        SAP();
        SAP(const std::vector<std::pair<size_t, size_t>> & _coords);
        // tokens "irreducible,index" (1-based) of one definition line
        SAP(const std::vector<std::string> & strs);
        ~SAP();

        const std::vector<std::pair<size_t, size_t>> & coords() const;
        const std::vector<std::pair<std::pair<size_t, size_t>, size_t>> & uniques_orders() const;
        const std::pair<size_t, size_t> & operator[](const size_t & index) const;

        size_t order() const;
        void pretty_print(std::ostream & stream) const;

        at::Tensor operator()(const std::vector<at::Tensor> & xs) const;
        std::vector<at::Tensor> gradient(const std::vector<at::Tensor> & xs) const;
        std::vector<at::Tensor> gradient_(const std::vector<at::Tensor> & xs) const;
        std::vector<at::Tensor> gradient_(const std::vector<at::Tensor> & xs, at::Tensor & grad) const;
        // ddP / dx^2: blocks hesses[i][j] (j >= i), upper triangle only (source/polynomial/SAP.cpp:177-322);
        // `hess` harvests the concatenated [sumdim][sumdim] matrix the blocks are views of
        CL::utility::matrix<at::Tensor> Hessian(const std::vector<at::Tensor> & xs) const;
        CL::utility::matrix<at::Tensor> Hessian_(const std::vector<at::Tensor> & xs) const;
        CL::utility::matrix<at::Tensor> Hessian_(const std::vector<at::Tensor> & xs, at::Tensor & hess) const;
};

} }

#endif
