// TEST INFRASTRUCTURE ONLY - never linked into or called from the product path.
//
// extern "C" driver around the UNMODIFIED reference classes (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libtchem_ref.so). The
// reference accepts one geometry per call (source/IC/IntCoordSet.cpp:140-141),
// so every entry point here loops over the batch on the host.
//
// Used by: tests/ (parity checker, golden-vector generation) and bench.py's
// cpu_baseline / --impl reference legs.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <torch/torch.h>

#include <tchem/IC/IntCoordSet.hpp>
#include <tchem/polynomial/SAPSet.hpp>
#include <tchem/polynomial/PolynomialSet.hpp>

namespace {
thread_local std::string g_error;
const c10::TensorOptions f64 = at::TensorOptions().dtype(torch::kFloat64);

at::Tensor view(const double * p, std::initializer_list<int64_t> sizes) {
    return at::from_blob(const_cast<double *>(p), at::IntArrayRef(sizes), f64);
}
void store(double * dst, const at::Tensor & src) {
    at::Tensor c = src.contiguous();
    std::memcpy(dst, c.data_ptr<double>(), sizeof(double) * c.numel());
}
template <typename F> int guarded(F && f) {
    try { f(); return 0; }
    catch (const std::exception & e) { g_error = e.what(); return 1; }
    catch (...) { g_error = "unknown exception"; return 2; }
}
using tchem::IC::IntCoordSet;
using tchem::polynomial::SAPSet;
struct SapHandle {
    SAPSet set;
    std::vector<int64_t> dims;
    int64_t sumdim;
};
std::vector<at::Tensor> split_xs(const SapHandle & h, const double * x) {
    std::vector<at::Tensor> xs(h.dims.size());
    int64_t off = 0;
    for (size_t i = 0; i < h.dims.size(); i++) {
        xs[i] = view(x + off, {h.dims[i]}).clone();
        off += h.dims[i];
    }
    return xs;
}
}

extern "C" {

const char * ref_last_error() {return g_error.c_str();}

void ref_set_num_threads(int n) {at::set_num_threads(n);}

void * ref_ic_create(const char * format, const char * file) {
    IntCoordSet * set = nullptr;
    int rc = guarded([&]{ set = new IntCoordSet(format, file); });
    return rc == 0 ? set : nullptr;
}
void ref_ic_destroy(void * h) {delete static_cast<IntCoordSet *>(h);}
int ref_ic_intdim(void * h) {return (int)static_cast<IntCoordSet *>(h)->size();}

// flattened definition table: one row per InvDisp = (ic index, coeff, type, natoms, atoms[5], min)
int ref_ic_nterms(void * h) {
    int n = 0;
    for (const auto & ic : static_cast<IntCoordSet *>(h)->intcoords()) n += (int)ic.coeff_invdisps().size();
    return n;
}
void ref_ic_terms(void * h, int * ic_index, double * coeff, int * type, int * natoms, int * atoms, double * min) {
    int t = 0, i = 0;
    for (const auto & ic : static_cast<IntCoordSet *>(h)->intcoords()) {
        for (const auto & cd : ic.coeff_invdisps()) {
            ic_index[t] = i; coeff[t] = cd.first; type[t] = (int)cd.second.type();
            natoms[t] = (int)cd.second.atoms().size();
            for (int k = 0; k < 5; k++) atoms[5 * t + k] = k < natoms[t] ? (int)cd.second.atoms()[k] : -1;
            min[t] = cd.second.min();
            t++;
        }
        i++;
    }
}

int ref_ic_q(void * h, const double * r, int64_t B, int64_t cartdim, double * q) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) store(q + b * n, set(view(r + b * cartdim, {cartdim})));
    });
}
int ref_ic_qJ(void * h, const double * r, int64_t B, int64_t cartdim, double * q, double * J) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) {
            at::Tensor qb, Jb;
            std::tie(qb, Jb) = set.compute_IC_J(view(r + b * cartdim, {cartdim}));
            store(q + b * n, qb);
            store(J + b * n * cartdim, Jb);
        }
    });
}
int ref_ic_qJK(void * h, const double * r, int64_t B, int64_t cartdim, double * q, double * J, double * K) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) {
            at::Tensor qb, Jb, Kb;
            // clone: compute_IC_J_K marks views of r as requiring grad (source/IC/InvDisp.cpp:649-650)
            std::tie(qb, Jb, Kb) = set.compute_IC_J_K(view(r + b * cartdim, {cartdim}).clone());
            store(q + b * n, qb);
            store(J + b * n * cartdim, Jb);
            store(K + b * n * cartdim * cartdim, Kb);
        }
    });
}
int ref_ic_grad_int2cart(void * h, const double * r, const double * intgrad, int64_t B, int64_t cartdim, double * cartgrad) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++)
        store(cartgrad + b * cartdim, set.gradient_int2cart(view(r + b * cartdim, {cartdim}), view(intgrad + b * n, {n})));
    });
}
int ref_ic_grad_cart2int(void * h, const double * r, const double * cartgrad, int64_t B, int64_t cartdim, double * intgrad) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++)
        store(intgrad + b * n, set.gradient_cart2int(view(r + b * cartdim, {cartdim}), view(cartgrad + b * cartdim, {cartdim})));
    });
}
int ref_ic_grad_cart2int_matrix(void * h, const double * r, int64_t B, int64_t cartdim, double * M) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++)
        store(M + b * n * cartdim, set.gradient_cart2int_matrix(view(r + b * cartdim, {cartdim})));
    });
}
int ref_ic_hess_int2cart(void * h, const double * r, const double * intgrad, const double * intHess,
                         int64_t B, int64_t cartdim, double * cartHess) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++)
        store(cartHess + b * cartdim * cartdim,
              set.Hessian_int2cart(view(r + b * cartdim, {cartdim}).clone(), view(intgrad + b * n, {n}),
                                   view(intHess + b * n * n, {n, n})));
    });
}
int ref_ic_hess_cart2int(void * h, const double * r, const double * cartgrad, const double * cartHess,
                         int64_t B, int64_t cartdim, double * intHess) {
    const IntCoordSet & set = *static_cast<IntCoordSet *>(h);
    int64_t n = set.size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++)
        store(intHess + b * n * n,
              set.Hessian_cart2int(view(r + b * cartdim, {cartdim}).clone(), view(cartgrad + b * cartdim, {cartdim}),
                                   view(cartHess + b * cartdim * cartdim, {cartdim, cartdim})));
    });
}

void * ref_sap_create(const char * file, int irreducible, const int64_t * dims, int ndims) {
    SapHandle * h = nullptr;
    int rc = guarded([&]{
        std::vector<size_t> d(dims, dims + ndims);
        h = new SapHandle{SAPSet(file, (size_t)irreducible, d), std::vector<int64_t>(dims, dims + ndims), 0};
        for (int i = 0; i < ndims; i++) h->sumdim += dims[i];
    });
    return rc == 0 ? h : nullptr;
}
void ref_sap_destroy(void * h) {delete static_cast<SapHandle *>(h);}
int ref_sap_nterms(void * h) {return (int)static_cast<SapHandle *>(h)->set.SAPs().size();}
// x: [B, sum(dims)] (irreducibles concatenated) -> y: [B, nterms]   (SAPSet::operator())
int ref_sap_value(void * hh, const double * x, int64_t B, double * y) {
    const SapHandle & h = *static_cast<SapHandle *>(hh);
    int64_t n = h.set.SAPs().size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) store(y + b * n, h.set(split_xs(h, x + b * h.sumdim)));
    });
}
// -> J: [B, nterms, sum(dims)]   (SAPSet::Jacobian_(xs, J), the concatenated form)
int ref_sap_jacobian(void * hh, const double * x, int64_t B, double * J) {
    const SapHandle & h = *static_cast<SapHandle *>(hh);
    int64_t n = h.set.SAPs().size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) {
            at::Tensor Jb;
            h.set.Jacobian_(split_xs(h, x + b * h.sumdim), Jb);
            store(J + b * n * h.sumdim, Jb);
        }
    });
}
// autograd-friendly variant SAPSet::Jacobian(xs), concatenated by the driver -> [B, nterms, sum(dims)]
int ref_sap_jacobian_pow(void * hh, const double * x, int64_t B, double * J) {
    const SapHandle & h = *static_cast<SapHandle *>(hh);
    int64_t n = h.set.SAPs().size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) {
            std::vector<at::Tensor> Js = h.set.Jacobian(split_xs(h, x + b * h.sumdim));
            store(J + b * n * h.sumdim, at::cat(Js, 1));
        }
    });
}
// -> K: [B, nterms, sum(dims), sum(dims)]   (SAPSet::Jacobian2nd_(xs, K), the concatenated symmetric form)
int ref_sap_jacobian2nd(void * hh, const double * x, int64_t B, double * K) {
    const SapHandle & h = *static_cast<SapHandle *>(hh);
    int64_t n = h.set.SAPs().size();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) {
            at::Tensor Kb;
            h.set.Jacobian2nd_(split_xs(h, x + b * h.sumdim), Kb);
            store(K + b * n * h.sumdim * h.sumdim, Kb);
        }
    });
}

// ---- PolynomialSet (source/polynomial/PolynomialSet.cpp): all terms up to `order` in `dimension` coordinates ----
void * ref_poly_create(int dimension, int order) {
    tchem::polynomial::PolynomialSet * h = nullptr;
    int rc = guarded([&]{ h = new tchem::polynomial::PolynomialSet((size_t)dimension, (size_t)order); });
    return rc == 0 ? h : nullptr;
}
void ref_poly_destroy(void * h) {delete static_cast<tchem::polynomial::PolynomialSet *>(h);}
int ref_poly_nterms(void * h) {return (int)static_cast<tchem::polynomial::PolynomialSet *>(h)->polynomials().size();}
// coords of every term, flattened; term_ptr[nterms + 1]. Returns the number of coords (call with NULLs to size).
int ref_poly_terms(void * hh, int32_t * term_ptr, int32_t * coords) {
    const auto & ps = static_cast<tchem::polynomial::PolynomialSet *>(hh)->polynomials();
    int n = 0;
    for (size_t i = 0; i < ps.size(); i++) {
        if (term_ptr) term_ptr[i] = n;
        for (size_t c : ps[i].coords()) { if (coords) coords[n] = (int32_t)c; n++; }
    }
    if (term_ptr) term_ptr[ps.size()] = n;
    return n;
}
// x: [B, dimension] -> y [B, nterms], J [B, nterms, dimension] (Jacobian_), K [B, nterms, dimension, dimension] (Jacobian2nd_)
int ref_poly_eval(void * hh, const double * x, int64_t B, double * y, double * J, double * K) {
    const auto & set = *static_cast<tchem::polynomial::PolynomialSet *>(hh);
    const int64_t n = set.polynomials().size(), d = set.dimension();
    return guarded([&]{
        for (int64_t b = 0; b < B; b++) {
            at::Tensor xb = view(x + b * d, {d}).clone();
            if (y) store(y + b * n, set(xb));
            if (J) store(J + b * n * d, set.Jacobian_(xb));
            if (K) store(K + b * n * d * d, set.Jacobian2nd_(xb));
        }
    });
}

} // extern "C"
