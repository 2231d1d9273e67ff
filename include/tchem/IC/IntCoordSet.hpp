// tchem::IC::IntCoordSet - drop-in for include/tchem/IC/IntCoordSet.hpp of Torch-Chemistry.
//
// Same constructors, accessors and evaluators. Differences, all supersets:
//  * r may be a batch r[B][cartdim]; every result then carries the leading B. 1-D input keeps
//    the reference's shapes.
//  * r may live on a CUDA device (results stay there, computed on the current stream of that
//    device) or on the host (copies in and out happen inside the call).
//  * the arithmetic runs in the CUDA library behind include/tchem_b200.h; there is no CPU path.
#ifndef tchem_IC_IntCoordSet_hpp
#define tchem_IC_IntCoordSet_hpp

#include <tchem/IC/IntCoord.hpp>

namespace tchem { namespace IC {

class IntCoordSet {
    private:
        std::vector<IntCoord> intcoords_;
        // compiled device tables, built on first use per (device, number of atoms)
        mutable std::shared_ptr<detail::Compiled> compiled_;
        detail::Compiled & compiled(const at::Tensor & r, const char * who) const;
    public:
        IntCoordSet();
        // file format (Columbus7, default), internal coordinate definition file
        IntCoordSet(const std::string & format, const std::string & file);
        // from already built coordinates
        IntCoordSet(const std::vector<IntCoord> & _intcoords);
        ~IntCoordSet();

        const std::vector<IntCoord> & intcoords() const;
        size_t size() const;
        const IntCoord & operator[](const size_t & index) const;

        void print(std::ofstream & ofs, const std::string & format) const;

        at::Tensor operator()(const at::Tensor & r) const;
        std::tuple<at::Tensor, at::Tensor> compute_IC_J(const at::Tensor & r) const;
        std::tuple<at::Tensor, at::Tensor, at::Tensor> compute_IC_J_K(const at::Tensor & r) const;

        at::Tensor gradient_cart2int_matrix(const at::Tensor & r) const;
        at::Tensor gradient_cart2int(const at::Tensor & r, const at::Tensor & cartgrad) const;
        at::Tensor gradient_int2cart(const at::Tensor & r, const at::Tensor & intgrad) const;

        at::Tensor Hessian_cart2int(const at::Tensor & r, const at::Tensor & cartgrad, const at::Tensor & cartHess) const;
        at::Tensor Hessian_int2cart(const at::Tensor & r, const at::Tensor & intgrad, const at::Tensor & intHess) const;
};

} }

#endif
