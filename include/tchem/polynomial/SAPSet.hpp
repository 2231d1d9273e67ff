// tchem::polynomial::SAPSet - drop-in for include/tchem/polynomial/SAPSet.hpp of Torch-Chemistry:
// construction from a term file, value, the three Jacobian variants and the three second-order
// Jacobian variants, evaluated by the CUDA library (batched: xs[i] may be [B][dim_i]).
// rotation / translation are set-up time utilities outside the B200 hot path and are not provided.
#ifndef tchem_polynomial_SAPSet_hpp
#define tchem_polynomial_SAPSet_hpp

#include <map>
#include <memory>

#include <tchem/polynomial/SAP.hpp>

namespace tchem { namespace IC { class IntCoordSet; namespace detail { class Compiled; } } }

namespace tchem { namespace polynomial {

namespace detail { class CompiledSAP; }

class SAPSet {
    private:
        std::vector<SAP> SAPs_;
        size_t irreducible_ = 0;
        std::vector<size_t> dimensions_;
        size_t max_order_ = 0;
        std::vector<std::vector<const SAP *>> orders_;
        void construct_orders_();
        mutable std::shared_ptr<detail::CompiledSAP> compiled_;
        // device tables of the IntCoordSets this set has been chained behind
        mutable std::shared_ptr<std::map<const void *, std::shared_ptr<tchem::IC::detail::Compiled>>> ic_compiled_;
        detail::CompiledSAP & compiled(int device) const;
        at::Tensor run_(const std::vector<at::Tensor> & xs, bool want_y, bool want_J, const char * who, at::Tensor * Jcat) const;
    public:
        SAPSet();
        SAPSet(const std::string & sapoly_file, const size_t & _irreducible, const std::vector<size_t> & _dimensions);
        // from explicit monomials (not in the reference; used by PolynomialSet)
        SAPSet(const std::vector<SAP> & _SAPs, const size_t & _irreducible, const std::vector<size_t> & _dimensions);
        ~SAPSet();

        void insert_const();
        const std::vector<SAP> & SAPs() const;
        const SAP & operator[](const size_t & index) const;
        void pretty_print(std::ostream & stream) const;

        at::Tensor operator()(const std::vector<at::Tensor> & xs) const;
        std::vector<at::Tensor> Jacobian(const std::vector<at::Tensor> & xs) const;
        std::vector<at::Tensor> Jacobian_(const std::vector<at::Tensor> & xs) const;
        std::vector<at::Tensor> Jacobian_(const std::vector<at::Tensor> & xs, at::Tensor & J) const;
        // dd{P(x)} / dx^2: blocks Ks[i][j] (j >= i) of shape [NSAP][dim_i][dim_j]; the overload with `K`
        // harvests the concatenated symmetric tensor [NSAP][sumdim][sumdim] the blocks are views of
        // (reference source/polynomial/SAPSet.cpp:203-276)
        CL::utility::matrix<at::Tensor> Jacobian2nd(const std::vector<at::Tensor> & xs) const;
        CL::utility::matrix<at::Tensor> Jacobian2nd_(const std::vector<at::Tensor> & xs) const;
        CL::utility::matrix<at::Tensor> Jacobian2nd_(const std::vector<at::Tensor> & xs, at::Tensor & K) const;

        // chained path (not in the reference, where callers slice q by hand): r -> q by the
        // compute_IC_J formulas -> xs = q split by dimensions_ -> value and concatenated Jacobian
        // d SAP / d q in one kernel. Returns (y, J, q).
        std::tuple<at::Tensor, at::Tensor, at::Tensor> chained(const tchem::IC::IntCoordSet & set, const at::Tensor & r) const;
        // chained path with CARTESIAN derivatives (what callers compose by hand from Jacobian_ / Jacobian2nd_ and
        // compute_IC_J / compute_IC_J_K): values, (d SAP / d q) J and - second_order - J^T (d2 SAP / d q2) J + sum_k (d SAP / d q_k) K_k
        // (an undefined tensor otherwise)
        std::tuple<at::Tensor, at::Tensor, at::Tensor> chained_cart(const tchem::IC::IntCoordSet & set, const at::Tensor & r,
                                                                    const bool & second_order = false) const;
};

} }

#endif
