// tchem::IC::IntCoord - drop-in for include/tchem/IC/IntCoord.hpp of Torch-Chemistry.
#ifndef tchem_IC_IntCoord_hpp
#define tchem_IC_IntCoord_hpp

#include <tchem/IC/InvDisp.hpp>

namespace tchem { namespace IC {

// a linear combination of invariant displacements
class IntCoord {
    private:
        std::vector<std::pair<double, InvDisp>> coeff_invdisps_;
    public:
        IntCoord();
        IntCoord(const std::vector<std::pair<double, InvDisp>> & _coeff_invdisps);
        ~IntCoord();

        const std::vector<std::pair<double, InvDisp>> & coeff_invdisps() const;
        const std::pair<double, InvDisp> & operator[](const size_t & index) const;

        void append(const std::pair<double, InvDisp> & coeff_invdisp);
        void append(const double & coeff, const InvDisp & invdisp);
        void normalize();

        void print(std::ofstream & ofs, const std::string & format) const;

        at::Tensor operator()(const at::Tensor & r) const;
        std::tuple<at::Tensor, at::Tensor> compute_IC_J(const at::Tensor & r) const;
        std::tuple<at::Tensor, at::Tensor, at::Tensor> compute_IC_J_K(const at::Tensor & r) const;
};

} }

#endif
