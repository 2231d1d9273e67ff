#ifndef tchem_chem_normal_mode_hpp
#define tchem_chem_normal_mode_hpp

#include <torch/torch.h>

// B200 drop-in for the reference's include/tchem/chem/normal_mode.hpp: same classes, constructors and accessors.
// Batched superset: Hessian / Jacobian may carry a leading batch axis (one analysis per geometry, all in one call).
// The eigen-solves run as batched device operations (cuSOLVER through ATen); host tensors are moved to the current
// CUDA device and the results come back on the host. There is no CPU path.
namespace tchem { namespace chem {

class CartNormalMode {
    protected:
        std::vector<double> masses_;
        at::Tensor Hessian_;
        bool ready_ = false;
        at::Tensor frequency_, cartmode_;
        bool host_ = false, batched_ = false;
        at::Tensor out(const at::Tensor & t) const;
    public:
        CartNormalMode();
        CartNormalMode(const std::vector<double> & _masses, const at::Tensor & _Hessian);
        ~CartNormalMode();

        // Perform normal mode analysis, then observables are ready for fetching
        void kernel();

        // Harmonic frequencies (negative if imaginary)
        at::Tensor frequency() const;
        // Cartesian coordinate normal modes in each row (normalized by mass metric)
        at::Tensor cartmode () const;
};

class IntNormalMode : public CartNormalMode {
    private:
        at::Tensor Jacobian_;
        at::Tensor intmode_, Linv_;
    public:
        IntNormalMode();
        IntNormalMode(const std::vector<double> & _masses, const at::Tensor & _Jacobian, const at::Tensor & _Hessian);
        ~IntNormalMode();

        void kernel();

        // Internal coordinate normal modes in each row (normalized by (J . M^-1 . J^T)^-1 metric)
        at::Tensor intmode() const;
        // L^-1 = (intmode_^T)^-1
        at::Tensor Linv   () const;
};

class SANormalMode {
    private:
        std::vector<IntNormalMode> parts_;
        bool ready_ = false;
    public:
        SANormalMode();
        SANormalMode(const std::vector<double> & _masses, const std::vector<at::Tensor> & _Jacobians, const std::vector<at::Tensor> & _Hessians);
        ~SANormalMode();

        void kernel();

        std::vector<at::Tensor> frequencies() const;
        std::vector<at::Tensor> intmodes   () const;
        std::vector<at::Tensor> Linvs      () const;
        std::vector<at::Tensor> cartmodes  () const;

        // Number of irreducible representations
        size_t NIrreds() const;
};

} // namespace chem
} // namespace tchem

#endif
