// tchem::IC::InvDisp - drop-in for Torch-Chemistry's include/tchem/IC/InvDisp.hpp (same enum,
// same class, same member signatures), evaluated by the B200 CUDA library behind tchem_b200.h.
// Every tensor argument may carry a leading batch axis: r[cartdim] or r[B][cartdim].
#ifndef tchem_IC_InvDisp_hpp
#define tchem_IC_InvDisp_hpp

#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include <torch/torch.h>

namespace tchem { namespace IC {

// same order and values as the reference (include/tchem/IC/InvDisp.hpp:9-44)
enum InvDisp_type {
    dummy, stretching, bending, cosbending, torsion, sintorsion, costorsion,
    torsion2, sintorsion2, costorsion2, pxtorsion2, pytorsion2, OutOfPlane, sinoop
};
static const std::vector<std::string> InvDisp_typestr = {
    "dummy", "stretching", "bending", "cosbending", "torsion", "sintorsion", "costorsion",
    "torsion2", "sintorsion2", "costorsion2", "pxtorsion2", "pytorsion2", "OutOfPlane", "sinoop"
};

namespace detail { class Compiled; }

class InvDisp {
    private:
        InvDisp_type type_;
        std::vector<size_t> atoms_;
        double min_ = -M_PI;
    public:
        InvDisp();
        InvDisp(const std::string & _type, const std::vector<size_t> & _atoms, const double & _min = -M_PI);
        ~InvDisp();

        const InvDisp_type & type() const;
        const std::vector<size_t> & atoms() const;
        const double & min() const;

        void print(std::ofstream & ofs, const std::string & format) const;

        at::Tensor operator()(const at::Tensor & r) const;
        std::tuple<at::Tensor, at::Tensor> compute_IC_J(const at::Tensor & r) const;
        std::tuple<at::Tensor, at::Tensor, at::Tensor> compute_IC_J_K(const at::Tensor & r) const;
};

} }

#endif
