// Host side of the drop-in: the tchem:: classes of include/tchem/ implemented over the C ABI of
// the CUDA library (include/tchem_b200.h). LibTorch is used for tensors, devices and streams
// only; every number is produced by the kernels in torch_chemistry_b200/csrc.
//
// Error convention of the reference: std::invalid_argument for shape
// errors with the fully qualified function name in the message, std::domain_error for
// unimplemented types, CL::utility::file_error for unreadable files; CUDA failures surface as
// std::runtime_error.
#include <map>
#include <mutex>
#include <thread>
#include <sstream>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <tchem/tchem.hpp>
#include <tchem/chem/normal_mode.hpp>
#include <tchem/utility.hpp>

#include "../include/tchem_b200.h"

namespace {

void check(int rc) {
    if (rc == TCHEM_OK) return;
    const std::string msg = tchem_last_error();
    switch (rc) {
        case TCHEM_EINVAL: throw std::invalid_argument(msg);
        case TCHEM_EDOMAIN: throw std::domain_error(msg);
        case TCHEM_EFILE: throw CL::utility::file_error(msg);
        default: throw std::runtime_error(msg);
    }
}

// Host tensors may be spread over several GPUs by ONE call (geometries are independent): the batch
// is cut into contiguous slices, one host thread and one table per device, each slice through the pipelined
// host-buffer entry point, results landing in place in the caller-visible tensors. No collective, no extra copy.
// TCHEM_DEVICES = "all" or a count (default 1); tchem::b200::set_host_devices() overrides the environment.
int g_host_devices = -1;
int host_devices() {
    int want = g_host_devices;
    if (want < 0) {
        const char * e = getenv("TCHEM_DEVICES");
        want = !e ? 1 : (std::string(e) == "all" ? 1 << 20 : atoi(e));
    }
    return std::max(1, std::min(want, tchem_device_count()));
}

int default_device() {
    int n = tchem_device_count();
    if (n <= 0) throw std::runtime_error("tchem: no CUDA device (this library has no CPU path)");
    return c10::cuda::current_device();
}

// validated view of a "vector or batch of vectors" argument
struct Batch {
    at::Tensor t;       // [B][width], contiguous float64
    bool batched;
    int64_t B, width;
};
Batch as_batch(const at::Tensor & x, const std::string & who, const std::string & what) {
    if (x.dim() != 1 && x.dim() != 2) throw std::invalid_argument(who + ": " + what + " must be a vector");
    if (x.scalar_type() != at::kDouble) throw std::invalid_argument(who + ": " + what + " must be float64");
    Batch b;
    b.batched = x.dim() == 2;
    b.t = (b.batched ? x : x.unsqueeze(0)).contiguous();
    b.B = b.t.size(0);
    b.width = b.t.size(1);
    return b;
}
at::Tensor unbatch(const at::Tensor & t, bool batched) { return batched ? t : t[0]; }
void * ptr(const at::Tensor & t) { return t.defined() && t.numel() > 0 ? t.data_ptr() : nullptr; }

} // namespace

namespace tchem { namespace b200 {
void set_host_devices(int n) { g_host_devices = n; }
int host_devices() { return ::host_devices(); }
} }

namespace tchem { namespace IC {

namespace detail {

// device tables of one IntCoordSet, one per (device, number of atoms)
class Compiled {
    std::vector<tchem_invdisp_t> terms_;
    int n_ic_;
    std::map<std::pair<int, int>, tchem_ic_table *> tables_;
    std::mutex mutex_;
public:
    Compiled(const std::vector<IntCoord> & ics) : n_ic_((int)ics.size()) {
        for (size_t i = 0; i < ics.size(); i++)
            for (const auto & cd : ics[i].coeff_invdisps()) {
                tchem_invdisp_t t;
                std::memset(&t, 0, sizeof(t));
                t.ic = (int)i;
                t.type = (int)cd.second.type();
                t.natoms = (int)cd.second.atoms().size();
                for (int a = 0; a < t.natoms && a < 5; a++) t.atoms[a] = (int)cd.second.atoms()[a];
                t.coeff = cd.first;
                t.min = cd.second.min();
                terms_.push_back(t);
            }
    }
    ~Compiled() { for (auto & kv : tables_) tchem_ic_table_destroy(kv.second); }
    tchem_ic_table * get(int device, int natoms) {
        std::lock_guard<std::mutex> lock(mutex_);
        auto key = std::make_pair(device, natoms);
        auto it = tables_.find(key);
        if (it != tables_.end()) return it->second;
        tchem_ic_table * t = nullptr;
        check(tchem_ic_table_create(terms_.data(), (int)terms_.size(), n_ic_, natoms, device, &t));
        tables_[key] = t;
        return t;
    }
};

} // namespace detail

namespace {

struct Call {                 // one evaluation request resolved against a table
    std::shared_ptr<detail::Compiled> compiled;
    tchem_ic_table * table;
    Batch r;
    bool host;
    int device;
    at::TensorOptions opts;   // where results live (same device as r)
    at::Tensor r_dev;         // r on the device (transforms)
};

Call resolve(const std::vector<IntCoord> & ics, std::shared_ptr<detail::Compiled> & compiled, const at::Tensor & r,
             const std::string & who) {
    Call c;
    c.r = as_batch(r, who, "r");
    if (c.r.width % 3 != 0) throw std::invalid_argument(who + ": r must hold 3 Cartesian coordinates per atom");
    if (!compiled) compiled = std::make_shared<detail::Compiled>(ics);
    c.host = !r.is_cuda();
    c.device = c.host ? default_device() : r.get_device();
    c.compiled = compiled;
    c.table = compiled->get(c.device, (int)(c.r.width / 3));
    c.opts = c.r.t.options();
    return c;
}

std::vector<at::Tensor> evaluate(Call & c, bool want_J, bool want_K, bool value_path) {
    const int64_t B = c.r.B, n = tchem_ic_table_intdim(c.table), cd = c.r.width;
    at::Tensor q = at::empty({B, n}, c.opts), J, K;
    if (want_J) J = at::empty({B, n, cd}, c.opts);
    if (want_K) K = at::empty({B, n, cd, cd}, c.opts);
    if (B > 0) {
        const int ndev = c.host ? host_devices() : 1;
        if (c.host && ndev > 1 && B >= 2048 * (int64_t)ndev) {
            // one slice, one table and one host thread per device; slices are multiples of the 32-geometry tile
            std::vector<std::thread> workers;
            std::vector<int> rcs(ndev, TCHEM_OK);
            std::vector<std::string> errs(ndev);
            const int64_t per = ((B + ndev - 1) / ndev + 31) / 32 * 32;
            const double * rp = (const double *)ptr(c.r.t);
            double * qp = (double *)ptr(q), * Jp = (double *)ptr(J), * Kp = (double *)ptr(K);
            for (int d = 0; d < ndev; d++) {
                const int64_t b0 = std::min<int64_t>(B, d * per), nb = std::min<int64_t>(per, B - b0);
                if (nb <= 0) continue;
                tchem_ic_table * table = c.compiled->get(d, (int)(cd / 3));
                workers.emplace_back([=, &rcs, &errs] {
                    rcs[d] = tchem_ic_eval_host(table, rp + b0 * cd, nb, qp ? qp + b0 * n : nullptr, Jp ? Jp + b0 * n * cd : nullptr,
                                                Kp ? Kp + b0 * n * cd * cd : nullptr, value_path ? 1 : 0);
                    if (rcs[d] != TCHEM_OK) errs[d] = tchem_last_error();
                });
            }
            for (auto & w : workers) w.join();
            for (int d = 0; d < ndev; d++)
                if (rcs[d] != TCHEM_OK) throw std::runtime_error(errs[d]);
        } else if (c.host) {
            check(tchem_ic_eval_host(c.table, (const double *)ptr(c.r.t), B, (double *)ptr(q), (double *)ptr(J),
                                     (double *)ptr(K), value_path ? 1 : 0));
        } else {
            c10::cuda::CUDAGuard guard(c.device);
            check(tchem_ic_eval(c.table, (const double *)ptr(c.r.t), B, (double *)ptr(q), (double *)ptr(J),
                                (double *)ptr(K), value_path ? 1 : 0, c10::cuda::getCurrentCUDAStream(c.device).stream()));
        }
    }
    std::vector<at::Tensor> out = {unbatch(q, c.r.batched)};
    if (want_J) out.push_back(unbatch(J, c.r.batched));
    if (want_K) out.push_back(unbatch(K, c.r.batched));
    return out;
}

at::Tensor on_device(const at::Tensor & t, int device) {
    return t.is_cuda() ? t : t.to(at::Device(at::kCUDA, device));
}

} // namespace

// ---- InvDisp ---------------------------------------------------------------------------------
InvDisp::InvDisp() : type_(dummy) {}
InvDisp::InvDisp(const std::string & _type, const std::vector<size_t> & _atoms, const double & _min)
: atoms_(_atoms), min_(_min) {
    size_t i = 0;
    for (; i < InvDisp_typestr.size(); i++) if (_type == InvDisp_typestr[i]) break;
    if (i == InvDisp_typestr.size()) throw std::invalid_argument("Unimplemented internal coordinate type " + _type);
    type_ = static_cast<InvDisp_type>(i);
}
InvDisp::~InvDisp() {}
const InvDisp_type & InvDisp::type() const { return type_; }
const std::vector<size_t> & InvDisp::atoms() const { return atoms_; }
const double & InvDisp::min() const { return min_; }

void InvDisp::print(std::ofstream & ofs, const std::string & format) const {
    if (format == "Columbus7") {
        // fixed-width "7." style atom fields; BEND lists the vertex last and OUT the centre last
        std::vector<size_t> order;
        std::string key;
        switch (type_) {
            case stretching: key = "STRE"; order = {0, 1}; break;
            case bending:    key = "BEND"; order = {0, 2, 1}; break;
            case torsion:    key = "TORS"; order = {0, 1, 2, 3}; break;
            case OutOfPlane: key = "OUT "; order = {0, 2, 3, 1}; break;
            default: throw std::invalid_argument("Columbus does not support " + InvDisp_typestr[type_]);
        }
        ofs << key;
        for (size_t k = 0; k < order.size(); k++)
            ofs << std::setw(k == 0 && type_ != stretching ? 10 : 9) << atoms_[order[k]] + 1 << '.';
    } else {
        ofs << std::setw(14) << InvDisp_typestr[type_];
        for (const size_t & atom : atoms_) ofs << std::setw(6) << atom + 1;
    }
    ofs << '\n';
}

namespace {
// a lone InvDisp / IntCoord is evaluated as a one-coordinate set
template <typename F> auto as_single_set(const std::vector<std::pair<double, InvDisp>> & terms, F && f) {
    IntCoordSet set(std::vector<IntCoord>{IntCoord(terms)});
    return f(set);
}
at::Tensor drop_ic_axis(const at::Tensor & t, bool batched) { return batched ? t.select(1, 0) : t.select(0, 0); }
}

at::Tensor InvDisp::operator()(const at::Tensor & r) const {
    if (r.dim() != 1 && r.dim() != 2) throw std::invalid_argument("tchem::IC::InvDisp::operator(): r must be a vector");
    return as_single_set({{1.0, *this}}, [&](IntCoordSet & s) { return drop_ic_axis(s(r), r.dim() == 2); });
}
std::tuple<at::Tensor, at::Tensor> InvDisp::compute_IC_J(const at::Tensor & r) const {
    if (r.dim() != 1 && r.dim() != 2) throw std::invalid_argument("tchem::IC::InvDisp::compute_IC_J: r must be a vector");
    return as_single_set({{1.0, *this}}, [&](IntCoordSet & s) {
        at::Tensor q, J;
        std::tie(q, J) = s.compute_IC_J(r);
        return std::make_tuple(drop_ic_axis(q, r.dim() == 2), drop_ic_axis(J, r.dim() == 2));
    });
}
std::tuple<at::Tensor, at::Tensor, at::Tensor> InvDisp::compute_IC_J_K(const at::Tensor & r) const {
    if (r.dim() != 1 && r.dim() != 2) throw std::invalid_argument("tchem::IC::InvDisp::compute_IC_J_K: r must be a vector");
    return as_single_set({{1.0, *this}}, [&](IntCoordSet & s) {
        at::Tensor q, J, K;
        std::tie(q, J, K) = s.compute_IC_J_K(r);
        const bool b = r.dim() == 2;
        return std::make_tuple(drop_ic_axis(q, b), drop_ic_axis(J, b), drop_ic_axis(K, b));
    });
}

// ---- IntCoord --------------------------------------------------------------------------------
IntCoord::IntCoord() {}
IntCoord::IntCoord(const std::vector<std::pair<double, InvDisp>> & _coeff_invdisps) : coeff_invdisps_(_coeff_invdisps) {}
IntCoord::~IntCoord() {}
const std::vector<std::pair<double, InvDisp>> & IntCoord::coeff_invdisps() const { return coeff_invdisps_; }
const std::pair<double, InvDisp> & IntCoord::operator[](const size_t & index) const { return coeff_invdisps_[index]; }
void IntCoord::append(const std::pair<double, InvDisp> & coeff_invdisp) { coeff_invdisps_.push_back(coeff_invdisp); }
void IntCoord::append(const double & coeff, const InvDisp & invdisp) { coeff_invdisps_.emplace_back(coeff, invdisp); }
void IntCoord::normalize() {
    double norm2 = 0.0;
    for (const auto & cd : coeff_invdisps_) norm2 += cd.first * cd.first;
    const double norm = norm2 / std::sqrt(norm2);
    for (auto & cd : coeff_invdisps_) cd.first /= norm;
}
void IntCoord::print(std::ofstream & ofs, const std::string & format) const {
    for (size_t i = 0; i < coeff_invdisps_.size(); i++) {
        if (format == "Columbus7")
            ofs << (i == 0 ? "K         " : "          ") << std::setw(9) << std::setprecision(6) << coeff_invdisps_[i].first << ' ';
        else
            ofs << (i == 0 ? "coord " : "      ") << std::setw(12) << coeff_invdisps_[i].first;
        coeff_invdisps_[i].second.print(ofs, format);
    }
}
at::Tensor IntCoord::operator()(const at::Tensor & r) const {
    if (r.dim() != 1 && r.dim() != 2) throw std::invalid_argument("tchem::IC::IntCoord::operator(): r must be a vector");
    return as_single_set(coeff_invdisps_, [&](IntCoordSet & s) { return drop_ic_axis(s(r), r.dim() == 2); });
}
std::tuple<at::Tensor, at::Tensor> IntCoord::compute_IC_J(const at::Tensor & r) const {
    if (r.dim() != 1 && r.dim() != 2) throw std::invalid_argument("tchem::IC::IntCoord::compute_IC_J: r must be a vector");
    return as_single_set(coeff_invdisps_, [&](IntCoordSet & s) {
        at::Tensor q, J;
        std::tie(q, J) = s.compute_IC_J(r);
        return std::make_tuple(drop_ic_axis(q, r.dim() == 2), drop_ic_axis(J, r.dim() == 2));
    });
}
std::tuple<at::Tensor, at::Tensor, at::Tensor> IntCoord::compute_IC_J_K(const at::Tensor & r) const {
    if (r.dim() != 1 && r.dim() != 2) throw std::invalid_argument("tchem::IC::IntCoord::compute_IC_J_K: r must be a vector");
    return as_single_set(coeff_invdisps_, [&](IntCoordSet & s) {
        at::Tensor q, J, K;
        std::tie(q, J, K) = s.compute_IC_J_K(r);
        const bool b = r.dim() == 2;
        return std::make_tuple(drop_ic_axis(q, b), drop_ic_axis(J, b), drop_ic_axis(K, b));
    });
}

// ---- IntCoordSet -----------------------------------------------------------------------------
IntCoordSet::IntCoordSet() {}
IntCoordSet::IntCoordSet(const std::vector<IntCoord> & _intcoords) : intcoords_(_intcoords) {}
IntCoordSet::IntCoordSet(const std::string & format, const std::string & file) {
    int n_terms = 0, n_ic = 0;
    check(tchem_ic_parse_file(format.c_str(), file.c_str(), nullptr, 0, &n_terms, &n_ic));
    std::vector<tchem_invdisp_t> terms(std::max(n_terms, 1));
    check(tchem_ic_parse_file(format.c_str(), file.c_str(), terms.data(), n_terms, &n_terms, &n_ic));
    intcoords_.resize(n_ic);
    for (int k = 0; k < n_terms; k++) {
        const tchem_invdisp_t & t = terms[k];
        std::vector<size_t> atoms(t.atoms, t.atoms + t.natoms);
        intcoords_[t.ic].append(t.coeff, InvDisp(InvDisp_typestr[t.type], atoms, t.min));   // already normalised
    }
}
IntCoordSet::~IntCoordSet() {}

const std::vector<IntCoord> & IntCoordSet::intcoords() const { return intcoords_; }
size_t IntCoordSet::size() const { return intcoords_.size(); }
const IntCoord & IntCoordSet::operator[](const size_t & index) const { return intcoords_[index]; }
void IntCoordSet::print(std::ofstream & ofs, const std::string & format) const {
    if (format == "Columbus7") ofs << "TEXAS\n";
    for (const IntCoord & ic : intcoords_) ic.print(ofs, format);
}

at::Tensor IntCoordSet::operator()(const at::Tensor & r) const {
    Call c = resolve(intcoords_, compiled_, r, "tchem::IC::IntCoordSet::operator()");
    return evaluate(c, false, false, true)[0];
}
std::tuple<at::Tensor, at::Tensor> IntCoordSet::compute_IC_J(const at::Tensor & r) const {
    Call c = resolve(intcoords_, compiled_, r, "tchem::IC::IntCoordSet::compute_IC_J");
    auto o = evaluate(c, true, false, false);
    return std::make_tuple(o[0], o[1]);
}
std::tuple<at::Tensor, at::Tensor, at::Tensor> IntCoordSet::compute_IC_J_K(const at::Tensor & r) const {
    Call c = resolve(intcoords_, compiled_, r, "tchem::IC::IntCoordSet::compute_IC_J_K");
    auto o = evaluate(c, true, true, false);
    return std::make_tuple(o[0], o[1], o[2]);
}

namespace {
// shared plumbing of the four transforms: device copies of the operands, launch, result back where r lives
// factors: the call factors J J^T per geometry - the reference's at::cholesky throws c10::Error when it is not
// positive definite (source/IC/IntCoordSet.cpp:182), so the device flag is read back (one synchronisation) and
// turned into the same exception type. TCHEM_ASYNC_TRANSFORMS=1 skips the check and keeps the call asynchronous;
// the affected geometries' results are NaN either way.
template <typename Launch>
at::Tensor transform(Call & c, std::vector<std::pair<at::Tensor, std::vector<int64_t>>> operands,
                     std::vector<int64_t> out_shape, Launch && launch, const char * factors = nullptr) {
    const int64_t B = c.r.B;
    c10::cuda::CUDAGuard guard(c.device);
    at::Tensor r_dev = on_device(c.r.t, c.device);
    std::vector<at::Tensor> ops;
    for (auto & op : operands) {
        std::vector<int64_t> shape = {B};
        shape.insert(shape.end(), op.second.begin(), op.second.end());
        ops.push_back(on_device(op.first.reshape(shape).contiguous(), c.device));
    }
    std::vector<int64_t> shape = {B};
    shape.insert(shape.end(), out_shape.begin(), out_shape.end());
    at::Tensor out = at::empty(shape, r_dev.options());
    if (B > 0) check(launch(r_dev, ops, out, c10::cuda::getCurrentCUDAStream(c.device).stream()));
    if (B > 0 && factors) {
        static const bool async = [] { const char * e = getenv("TCHEM_ASYNC_TRANSFORMS"); return e && atoi(e) != 0; }();
        if (!async && tchem_ic_factor_status(c.table) != 0)
            TORCH_CHECK(false, factors, ": cholesky: J J^T is not positive definite for at least one geometry of the batch "
                                        "(redundant or singular internal coordinates)");
    }
    if (c.host) out = out.cpu();
    return unbatch(out, c.r.batched);
}
void expect(bool ok, const std::string & msg) { if (!ok) throw std::invalid_argument(msg); }
}

at::Tensor IntCoordSet::gradient_cart2int_matrix(const at::Tensor & r) const {
    Call c = resolve(intcoords_, compiled_, r, "tchem::IC::IntCoordSet::gradient_cart2int");
    const int64_t n = size(), cd = c.r.width;
    return transform(c, {}, {n, cd}, [&](const at::Tensor & rd, std::vector<at::Tensor> &, at::Tensor & out, cudaStream_t s) {
        return tchem_ic_grad_cart2int_matrix(c.table, (const double *)ptr(rd), c.r.B, (double *)ptr(out), s);
    }, "tchem::IC::IntCoordSet::gradient_cart2int_matrix");
}
at::Tensor IntCoordSet::gradient_cart2int(const at::Tensor & r, const at::Tensor & cartgrad) const {
    const std::string who = "tchem::IC::IntCoordSet::gradient_cart2int";
    Call c = resolve(intcoords_, compiled_, r, who);
    expect(cartgrad.dim() == r.dim(), who + ": Cartesian coordinate gradient must be a vector");
    expect(cartgrad.size(-1) == c.r.width, who + ": Cartesian coordinate gradient must share a same dimension with r");
    const int64_t n = size(), cd = c.r.width;
    return transform(c, {{cartgrad, {cd}}}, {n}, [&](const at::Tensor & rd, std::vector<at::Tensor> & o, at::Tensor & out, cudaStream_t s) {
        return tchem_ic_grad_cart2int(c.table, (const double *)ptr(rd), (const double *)ptr(o[0]), c.r.B, (double *)ptr(out), s);
    }, "tchem::IC::IntCoordSet::gradient_cart2int");
}
at::Tensor IntCoordSet::gradient_int2cart(const at::Tensor & r, const at::Tensor & intgrad) const {
    const std::string who = "tchem::IC::IntCoordSet::gradient_int2cart";
    Call c = resolve(intcoords_, compiled_, r, who);
    expect(intgrad.dim() == r.dim(), who + ": Internal coordinate gradient must be a vector");
    expect(intgrad.size(-1) == (int64_t)size(), who + ": Internal coordinate gradient must share a same dimension with q");
    const int64_t n = size(), cd = c.r.width;
    return transform(c, {{intgrad, {n}}}, {cd}, [&](const at::Tensor & rd, std::vector<at::Tensor> & o, at::Tensor & out, cudaStream_t s) {
        return tchem_ic_grad_int2cart(c.table, (const double *)ptr(rd), (const double *)ptr(o[0]), c.r.B, (double *)ptr(out), s);
    });
}
at::Tensor IntCoordSet::Hessian_cart2int(const at::Tensor & r, const at::Tensor & cartgrad, const at::Tensor & cartHess) const {
    const std::string who = "tchem::IC::IntCoordSet::Hessian_cart2int";
    Call c = resolve(intcoords_, compiled_, r, who);
    const int64_t n = size(), cd = c.r.width;
    expect(cartgrad.dim() == r.dim(), who + ": Cartesian coordinate gradient must be a vector");
    expect(cartgrad.size(-1) == cd, who + ": Cartesian coordinate gradient must share a same dimension with r");
    expect(cartHess.dim() == r.dim() + 1, who + ": Cartesian coordinate Hessian must be a matrix");
    expect(cartHess.size(-1) == cartHess.size(-2), who + ": Cartesian coordinate Hessian must be a square matrix");
    expect(cartHess.size(-1) == cd, who + ": The dimensions of the gradient and the Hessian must match");
    return transform(c, {{cartgrad, {cd}}, {cartHess, {cd, cd}}}, {n, n},
                     [&](const at::Tensor & rd, std::vector<at::Tensor> & o, at::Tensor & out, cudaStream_t s) {
        return tchem_ic_hess_cart2int(c.table, (const double *)ptr(rd), (const double *)ptr(o[0]), (const double *)ptr(o[1]),
                                      c.r.B, (double *)ptr(out), s);
    }, "tchem::IC::IntCoordSet::Hessian_cart2int");
}
at::Tensor IntCoordSet::Hessian_int2cart(const at::Tensor & r, const at::Tensor & intgrad, const at::Tensor & intHess) const {
    const std::string who = "tchem::IC::IntCoordSet::Hessian_int2cart";
    Call c = resolve(intcoords_, compiled_, r, who);
    const int64_t n = size(), cd = c.r.width;
    expect(intgrad.dim() == r.dim(), who + ": Internal coordinate gradient must be a vector");
    expect(intHess.dim() == r.dim() + 1, who + ": Internal coordinate Hessian must be a matrix");
    expect(intHess.size(-1) == intHess.size(-2), who + ": Internal coordinate Hessian must be a square matrix");
    expect(intgrad.size(-1) == intHess.size(-1), who + ": The dimensions of the gradient and the Hessian must match");
    expect(intgrad.size(-1) == n, who + ": Internal coordinate gradient must share a same dimension with q");
    return transform(c, {{intgrad, {n}}, {intHess, {n, n}}}, {cd, cd},
                     [&](const at::Tensor & rd, std::vector<at::Tensor> & o, at::Tensor & out, cudaStream_t s) {
        return tchem_ic_hess_int2cart(c.table, (const double *)ptr(rd), (const double *)ptr(o[0]), (const double *)ptr(o[1]),
                                      c.r.B, (double *)ptr(out), s);
    });
}

// used by SAPSet::chained
tchem_ic_table * table_of(const IntCoordSet & set, std::shared_ptr<detail::Compiled> & compiled, int device, int natoms) {
    if (!compiled) compiled = std::make_shared<detail::Compiled>(set.intcoords());
    return compiled->get(device, natoms);
}

} } // namespace tchem::IC

// ==============================================================================================
namespace tchem { namespace IC {
tchem_ic_table * table_of(const IntCoordSet & set, std::shared_ptr<detail::Compiled> & compiled, int device, int natoms);
} }

namespace tchem { namespace polynomial {

namespace detail {
class CompiledSAP {
    std::vector<int32_t> term_ptr_, coords_, dims_;
    int irreducible_;
    std::map<int, tchem_sap_table *> tables_;
    std::mutex mutex_;
public:
    CompiledSAP(const std::vector<SAP> & saps, size_t irreducible, const std::vector<size_t> & dims) : irreducible_((int)irreducible) {
        term_ptr_.push_back(0);
        for (const SAP & s : saps) {
            for (const auto & c : s.coords()) { coords_.push_back((int32_t)c.first); coords_.push_back((int32_t)c.second); }
            term_ptr_.push_back((int32_t)(coords_.size() / 2));
        }
        for (size_t d : dims) dims_.push_back((int32_t)d);
    }
    ~CompiledSAP() { for (auto & kv : tables_) tchem_sap_table_destroy(kv.second); }
    tchem_sap_table * get(int device) {
        std::lock_guard<std::mutex> lock(mutex_);
        auto it = tables_.find(device);
        if (it != tables_.end()) return it->second;
        tchem_sap_table * t = nullptr;
        check(tchem_sap_table_create(term_ptr_.data(), coords_.data(), (int)term_ptr_.size() - 1, irreducible_, dims_.data(),
                                     (int)dims_.size(), device, &t));
        tables_[device] = t;
        return t;
    }
};
}

// ---- SAP ---------------------------------------------------------------------------------------
SAP::SAP() {}
SAP::SAP(const std::vector<std::pair<size_t, size_t>> & _coords) : coords_(_coords) { build_uniques_(); }
SAP::SAP(const std::vector<std::string> & strs) {
    for (const std::string & s : strs) {
        const size_t comma = s.find(',');
        if (comma == std::string::npos) throw std::invalid_argument("tchem::polynomial::SAP: bad token " + s);
        coords_.emplace_back(std::stoul(s.substr(0, comma)) - 1, std::stoul(s.substr(comma + 1)) - 1);
    }
    build_uniques_();
}
SAP::~SAP() {}
void SAP::build_uniques_() {
    std::sort(coords_.begin(), coords_.end(), std::greater<std::pair<size_t, size_t>>());
    uniques_orders_.clear();
    std::map<std::pair<size_t, size_t>, size_t> count;
    for (const auto & c : coords_) count[c]++;
    for (const auto & kv : count) uniques_orders_.push_back({kv.first, kv.second});
}
const std::vector<std::pair<size_t, size_t>> & SAP::coords() const { return coords_; }
const std::vector<std::pair<std::pair<size_t, size_t>, size_t>> & SAP::uniques_orders() const { return uniques_orders_; }
const std::pair<size_t, size_t> & SAP::operator[](const size_t & index) const { return coords_[index]; }
size_t SAP::order() const { return coords_.size(); }
void SAP::pretty_print(std::ostream & stream) const {
    stream << coords_.size() << "    ";
    for (const auto & c : coords_) stream << c.first + 1 << ',' << c.second + 1 << "    ";
    stream << '\n';
}
at::Tensor SAP::operator()(const std::vector<at::Tensor> & xs) const {
    for (const at::Tensor & x : xs) if (x.dim() != 1) throw std::invalid_argument("tchem::polynomial::SAP::operator(): x must be a vector");
    at::Tensor value = xs[0].new_full({}, 1.0);
    for (const auto & c : coords_) value = value * xs[c.first][c.second];
    return value;
}
std::vector<at::Tensor> SAP::gradient(const std::vector<at::Tensor> & xs) const {
    for (const at::Tensor & x : xs) if (x.dim() != 1) throw std::invalid_argument("tchem::polynomial::SAP::gradient: x must be a vector");
    std::vector<at::Tensor> grads(xs.size());
    for (size_t i = 0; i < xs.size(); i++) grads[i] = xs[i].new_zeros(xs[i].sizes());
    for (size_t i = 0; i < uniques_orders_.size(); i++) {
        const auto & u = uniques_orders_[i];
        at::Tensor g = (double)u.second * at::pow(xs[u.first.first][u.first.second], (double)(u.second - 1));
        for (size_t j = 0; j < uniques_orders_.size(); j++) if (j != i) {
            const auto & v = uniques_orders_[j];
            g = g * at::pow(xs[v.first.first][v.first.second], (double)v.second);
        }
        grads[u.first.first][u.first.second] = g;
    }
    return grads;
}
std::vector<at::Tensor> SAP::gradient_(const std::vector<at::Tensor> & xs) const { return gradient(xs); }
std::vector<at::Tensor> SAP::gradient_(const std::vector<at::Tensor> & xs, at::Tensor & grad) const {
    std::vector<at::Tensor> parts = gradient(xs);
    grad = at::cat(parts);
    std::vector<at::Tensor> views;
    int64_t start = 0;
    for (const at::Tensor & x : xs) { views.push_back(grad.slice(0, start, start + x.size(0))); start += x.size(0); }
    return views;
}

// source/polynomial/SAP.cpp:177-322: diagonal p (p-1) x^(p-2) prod_{others}, strict upper triangle
// p_u p_v x_u^(p_u-1) x_v^(p_v-1) prod_{others}; the lower triangle is left zero as in the reference
CL::utility::matrix<at::Tensor> SAP::Hessian_(const std::vector<at::Tensor> & xs, at::Tensor & hess) const {
    for (const at::Tensor & x : xs) if (x.dim() != 1) throw std::invalid_argument("tchem::polynomial::SAP::Hessian_: x must be a vector");
    int64_t dimension = 0;
    for (const at::Tensor & x : xs) dimension += x.size(0);
    hess = xs[0].new_zeros({dimension, dimension});
    CL::utility::matrix<at::Tensor> hesses(xs.size());
    std::vector<int64_t> start(xs.size() + 1, 0);
    for (size_t i = 0; i < xs.size(); i++) start[i + 1] = start[i] + xs[i].size(0);
    for (size_t i = 0; i < xs.size(); i++)
    for (size_t j = i; j < xs.size(); j++)
    hesses[i][j] = hess.slice(0, start[i], start[i + 1]).slice(1, start[j], start[j + 1]);
    auto power = [&](size_t k, int64_t p) { return at::pow(xs[uniques_orders_[k].first.first][uniques_orders_[k].first.second], (double)p); };
    const size_t n = uniques_orders_.size();
    for (size_t i = 0; i < n; i++) {
        const auto & ui = uniques_orders_[i];
        const int64_t pi = (int64_t)ui.second;
        if (pi >= 2) {
            at::Tensor el = (double)(pi * (pi - 1)) * power(i, pi - 2);
            for (size_t k = 0; k < n; k++) if (k != i) el = el * power(k, (int64_t)uniques_orders_[k].second);
            hesses[ui.first.first][ui.first.first][ui.first.second][ui.first.second] = el;
        }
        for (size_t j = i + 1; j < n; j++) {
            const auto & uj = uniques_orders_[j];
            const int64_t pj = (int64_t)uj.second;
            at::Tensor el = (double)(pi * pj) * power(i, pi - 1) * power(j, pj - 1);
            for (size_t k = 0; k < n; k++) if (k != i && k != j) el = el * power(k, (int64_t)uniques_orders_[k].second);
            hesses[ui.first.first][uj.first.first][ui.first.second][uj.first.second] = el;
        }
    }
    return hesses;
}
CL::utility::matrix<at::Tensor> SAP::Hessian_(const std::vector<at::Tensor> & xs) const {
    at::Tensor hess;
    CL::utility::matrix<at::Tensor> views = Hessian_(xs, hess);
    for (size_t i = 0; i < xs.size(); i++)
    for (size_t j = i; j < xs.size(); j++) views[i][j] = views[i][j].clone();
    return views;
}
CL::utility::matrix<at::Tensor> SAP::Hessian(const std::vector<at::Tensor> & xs) const { return Hessian_(xs); }

// ---- SAPSet ------------------------------------------------------------------------------------
SAPSet::SAPSet() {}
SAPSet::SAPSet(const std::string & sapoly_file, const size_t & _irreducible, const std::vector<size_t> & _dimensions)
: irreducible_(_irreducible), dimensions_(_dimensions) {
    int nt = 0, nc = 0;
    check(tchem_sap_parse_file(sapoly_file.c_str(), nullptr, 0, nullptr, 0, &nt, &nc));
    std::vector<int32_t> tp(nt + 1), cs(2 * std::max(nc, 1));
    check(tchem_sap_parse_file(sapoly_file.c_str(), tp.data(), nt, cs.data(), nc, &nt, &nc));
    for (int t = 0; t < nt; t++) {
        std::vector<std::pair<size_t, size_t>> coords;
        for (int c = tp[t]; c < tp[t + 1]; c++) coords.emplace_back(cs[2 * c], cs[2 * c + 1]);
        SAPs_.push_back(SAP(coords));
    }
    construct_orders_();
}
SAPSet::SAPSet(const std::vector<SAP> & _SAPs, const size_t & _irreducible, const std::vector<size_t> & _dimensions)
: SAPs_(_SAPs), irreducible_(_irreducible), dimensions_(_dimensions) { construct_orders_(); }
SAPSet::~SAPSet() {}
void SAPSet::construct_orders_() {
    max_order_ = 0;
    for (const SAP & s : SAPs_) max_order_ = std::max(max_order_, s.order());
    orders_.assign(max_order_ + 1, {});
    for (const SAP & s : SAPs_) orders_[s.order()].push_back(&s);
    if (irreducible_ != 0 && !orders_[0].empty()) throw std::invalid_argument(
        "tchem::SAPSet::construct_orders_: Only the totally symmetric irreducible can have 0th order term");
    compiled_.reset();
}
void SAPSet::insert_const() {
    if (irreducible_ == 0 && orders_[0].empty()) {
        SAPs_.insert(SAPs_.begin(), SAP());
        construct_orders_();
    }
}
const std::vector<SAP> & SAPSet::SAPs() const { return SAPs_; }
const SAP & SAPSet::operator[](const size_t & index) const { return SAPs_[index]; }
void SAPSet::pretty_print(std::ostream & stream) const { for (const SAP & s : SAPs_) s.pretty_print(stream); }

detail::CompiledSAP & SAPSet::compiled(int) const {
    if (!compiled_) compiled_ = std::make_shared<detail::CompiledSAP>(SAPs_, irreducible_, dimensions_);
    return *compiled_;
}

namespace {
at::Tensor on_device_cat(const std::vector<at::Tensor> & parts, int device) {
    at::Tensor x = at::cat(parts, 1).to(at::kDouble);
    return (x.is_cuda() ? x : x.to(at::Device(at::kCUDA, device))).contiguous();
}
}

// shared by the four evaluators: argument checks of source/polynomial/SAPSet.cpp:135-140, one launch
at::Tensor SAPSet::run_(const std::vector<at::Tensor> & xs, bool want_y, bool want_J, const char * who_, at::Tensor * Jcat) const {
    const std::string who = std::string("tchem::polynomial::SAPSet::") + who_;
    for (const at::Tensor & x : xs) if (x.dim() != 1 && x.dim() != 2) throw std::invalid_argument(who + ": x must be a vector");
    if (xs.size() != dimensions_.size()) throw std::invalid_argument(
        who + ": x must share a same number of irreducibles as the coordinate system");
    for (size_t i = 0; i < xs.size(); i++) if (xs[i].size(-1) != (int64_t)dimensions_[i]) throw std::invalid_argument(
        who + ": x must have a same dimension as the coordinates");
    const bool batched = xs[0].dim() == 2, host = !xs[0].is_cuda();
    const int device = host ? default_device() : xs[0].get_device();
    std::vector<at::Tensor> parts;
    for (const at::Tensor & x : xs) parts.push_back(batched ? x : x.unsqueeze(0));
    at::Tensor x = on_device_cat(parts, device);
    const int64_t B = x.size(0), nsap = SAPs_.size(), sumdim = x.size(1);
    c10::cuda::CUDAGuard guard(device);
    at::Tensor y, J;
    if (want_y) y = at::empty({B, nsap}, x.options());
    if (want_J) J = at::empty({B, nsap, sumdim}, x.options());
    if (B > 0) check(tchem_sap_eval(compiled(device).get(device), (const double *)ptr(x), B, (double *)ptr(y), (double *)ptr(J),
                                    c10::cuda::getCurrentCUDAStream(device).stream()));
    if (host) { if (want_y) y = y.cpu(); if (want_J) J = J.cpu(); }
    if (want_J && Jcat) *Jcat = unbatch(J, batched);
    return want_y ? unbatch(y, batched) : at::Tensor();
}
at::Tensor SAPSet::operator()(const std::vector<at::Tensor> & xs) const { return run_(xs, true, false, "operator()", nullptr); }
std::vector<at::Tensor> SAPSet::Jacobian_(const std::vector<at::Tensor> & xs, at::Tensor & J) const {
    run_(xs, false, true, "Jacobian_", &J);
    std::vector<at::Tensor> views;
    int64_t start = 0;
    for (size_t d : dimensions_) { views.push_back(J.slice(-1, start, start + (int64_t)d)); start += (int64_t)d; }
    return views;
}
std::vector<at::Tensor> SAPSet::Jacobian_(const std::vector<at::Tensor> & xs) const {
    at::Tensor J;
    std::vector<at::Tensor> views = Jacobian_(xs, J);
    for (at::Tensor & v : views) v = v.contiguous();
    return views;
}
std::vector<at::Tensor> SAPSet::Jacobian(const std::vector<at::Tensor> & xs) const { return Jacobian_(xs); }

// second-order Jacobian (source/polynomial/SAPSet.cpp:203-276): one launch fills the concatenated
// symmetric K; the blocks are views of it
CL::utility::matrix<at::Tensor> SAPSet::Jacobian2nd_(const std::vector<at::Tensor> & xs, at::Tensor & K) const {
    const std::string who = "tchem::polynomial::SAPSet::Jacobian2nd_";
    for (const at::Tensor & x : xs) if (x.dim() != 1 && x.dim() != 2) throw std::invalid_argument(who + ": x must be a vector");
    if (xs.size() != dimensions_.size()) throw std::invalid_argument(
        who + ": x must share a same number of irreducibles as the coordinate system");
    for (size_t i = 0; i < xs.size(); i++) if (xs[i].size(-1) != (int64_t)dimensions_[i]) throw std::invalid_argument(
        who + ": x must have a same dimension as the coordinates");
    const bool batched = xs[0].dim() == 2, host = !xs[0].is_cuda();
    const int device = host ? default_device() : xs[0].get_device();
    std::vector<at::Tensor> parts;
    for (const at::Tensor & x : xs) parts.push_back(batched ? x : x.unsqueeze(0));
    at::Tensor x = on_device_cat(parts, device);
    const int64_t B = x.size(0), nsap = SAPs_.size(), sumdim = x.size(1);
    c10::cuda::CUDAGuard guard(device);
    at::Tensor Kd = at::empty({B, nsap, sumdim, sumdim}, x.options());
    if (B > 0) check(tchem_sap_jacobian2nd(compiled(device).get(device), (const double *)ptr(x), B, (double *)ptr(Kd),
                                           c10::cuda::getCurrentCUDAStream(device).stream()));
    if (host) Kd = Kd.cpu();
    K = unbatch(Kd, batched);
    CL::utility::matrix<at::Tensor> Ks(xs.size());
    int64_t start_row = 0;
    for (size_t i = 0; i < xs.size(); i++) {
        const int64_t stop_row = start_row + (int64_t)dimensions_[i];
        int64_t start_col = start_row;
        for (size_t j = i; j < xs.size(); j++) {
            const int64_t stop_col = start_col + (int64_t)dimensions_[j];
            Ks[i][j] = K.slice(-2, start_row, stop_row).slice(-1, start_col, stop_col);
            start_col = stop_col;
        }
        start_row = stop_row;
    }
    return Ks;
}
// without `K` the reference fills the upper triangle only (SAP.cpp:213-263 never writes the strict
// lower triangle of a diagonal block)
CL::utility::matrix<at::Tensor> SAPSet::Jacobian2nd_(const std::vector<at::Tensor> & xs) const {
    at::Tensor K;
    CL::utility::matrix<at::Tensor> Ks = Jacobian2nd_(xs, K);
    for (size_t i = 0; i < xs.size(); i++)
    for (size_t j = i; j < xs.size(); j++)
    Ks[i][j] = i == j ? at::triu(Ks[i][j]) : Ks[i][j].contiguous();
    return Ks;
}
CL::utility::matrix<at::Tensor> SAPSet::Jacobian2nd(const std::vector<at::Tensor> & xs) const { return Jacobian2nd_(xs); }

std::tuple<at::Tensor, at::Tensor, at::Tensor> SAPSet::chained(const tchem::IC::IntCoordSet & set, const at::Tensor & r) const {
    const std::string who = "tchem::polynomial::SAPSet::chained";
    Batch rb = as_batch(r, who, "r");
    const bool host = !r.is_cuda();
    const int device = host ? default_device() : r.get_device();
    int64_t sumdim = 0;
    for (size_t d : dimensions_) sumdim += (int64_t)d;
    if ((int64_t)set.size() != sumdim) throw std::invalid_argument(who + ": x must have a same dimension as the coordinates");
    c10::cuda::CUDAGuard guard(device);
    at::Tensor rd = rb.t.is_cuda() ? rb.t : rb.t.to(at::Device(at::kCUDA, device));
    const int64_t B = rb.B, nsap = SAPs_.size();
    at::Tensor y = at::empty({B, nsap}, rd.options()), J = at::empty({B, nsap, sumdim}, rd.options()), q = at::empty({B, sumdim}, rd.options());
    if (!ic_compiled_) ic_compiled_ = std::make_shared<std::map<const void *, std::shared_ptr<tchem::IC::detail::Compiled>>>();
    std::shared_ptr<tchem::IC::detail::Compiled> & slot = (*ic_compiled_)[&set];
    tchem_ic_table * ict = tchem::IC::table_of(set, slot, device, (int)(rb.width / 3));
    if (B > 0) check(tchem_ic_sap_eval(ict, compiled(device).get(device), (const double *)ptr(rd), B, (double *)ptr(q), (double *)ptr(y),
                                       (double *)ptr(J), c10::cuda::getCurrentCUDAStream(device).stream()));
    if (host) { y = y.cpu(); J = J.cpu(); q = q.cpu(); }
    return std::make_tuple(unbatch(y, rb.batched), unbatch(J, rb.batched), unbatch(q, rb.batched));
}
std::tuple<at::Tensor, at::Tensor, at::Tensor> SAPSet::chained_cart(const tchem::IC::IntCoordSet & set, const at::Tensor & r,
                                                                    const bool & second_order) const {
    const std::string who = "tchem::polynomial::SAPSet::chained_cart";
    Batch rb = as_batch(r, who, "r");
    const bool host = !r.is_cuda();
    const int device = host ? default_device() : r.get_device();
    int64_t sumdim = 0;
    for (size_t d : dimensions_) sumdim += (int64_t)d;
    if ((int64_t)set.size() != sumdim) throw std::invalid_argument(who + ": x must have a same dimension as the coordinates");
    c10::cuda::CUDAGuard guard(device);
    at::Tensor rd = rb.t.is_cuda() ? rb.t : rb.t.to(at::Device(at::kCUDA, device));
    const int64_t B = rb.B, nsap = SAPs_.size(), cd = rb.width;
    at::Tensor y = at::empty({B, nsap}, rd.options()), g = at::empty({B, nsap, cd}, rd.options()), h;
    if (second_order) h = at::empty({B, nsap, cd, cd}, rd.options());
    if (!ic_compiled_) ic_compiled_ = std::make_shared<std::map<const void *, std::shared_ptr<tchem::IC::detail::Compiled>>>();
    std::shared_ptr<tchem::IC::detail::Compiled> & slot = (*ic_compiled_)[&set];
    tchem_ic_table * ict = tchem::IC::table_of(set, slot, device, (int)(rb.width / 3));
    if (B > 0) check(tchem_ic_sap_eval_cart(ict, compiled(device).get(device), (const double *)ptr(rd), B, (double *)ptr(y), (double *)ptr(g),
                                            (double *)ptr(h), c10::cuda::getCurrentCUDAStream(device).stream()));
    if (host) { y = y.cpu(); g = g.cpu(); if (second_order) h = h.cpu(); }
    return std::make_tuple(unbatch(y, rb.batched), unbatch(g, rb.batched), second_order ? unbatch(h, rb.batched) : h);
}

// ---- Polynomial / PolynomialSet (source/polynomial/Polynomial.cpp, PolynomialSet.cpp:56-179) --------------------
Polynomial::Polynomial() {}
Polynomial::Polynomial(const std::vector<size_t> & _coords) : coords_(_coords) {
    std::sort(coords_.begin(), coords_.end(), std::greater<size_t>());
    std::map<size_t, size_t> powers;
    for (size_t c : coords_) powers[c]++;
    for (const auto & kv : powers) uniques_orders_.push_back(kv);
}
Polynomial::Polynomial(const std::vector<std::pair<size_t, size_t>> _uniques_orders) : uniques_orders_(_uniques_orders) {
    for (const auto & uo : uniques_orders_) for (size_t i = 0; i < uo.second; i++) coords_.push_back(uo.first);
    std::sort(coords_.begin(), coords_.end(), std::greater<size_t>());
}
Polynomial::~Polynomial() {}
const std::vector<size_t> & Polynomial::coords() const { return coords_; }
const std::vector<std::pair<size_t, size_t>> & Polynomial::uniques_orders() const { return uniques_orders_; }
size_t Polynomial::order() const { return coords_.size(); }
namespace {
PolynomialSet single_term(const Polynomial & p, const at::Tensor & x, const char * who) {
    if (x.dim() != 1 && x.dim() != 2) throw std::invalid_argument(std::string("tchem::polynomial::Polynomial::") + who + ": x must be a vector");
    return PolynomialSet(std::vector<Polynomial>{p}, (size_t)x.size(-1));
}
}
at::Tensor Polynomial::operator()(const at::Tensor & x) const { return single_term(*this, x, "operator()")(x).select(-1, 0); }
at::Tensor Polynomial::gradient_(const at::Tensor & x) const { return single_term(*this, x, "gradient_").Jacobian_(x).select(-2, 0); }
at::Tensor Polynomial::gradient(const at::Tensor & x) const { return gradient_(x); }
at::Tensor Polynomial::Hessian_(const at::Tensor & x) const { return single_term(*this, x, "Hessian_").Jacobian2nd_(x).select(-3, 0); }
at::Tensor Polynomial::Hessian(const at::Tensor & x) const { return Hessian_(x); }

void PolynomialSet::construct_orders_() {
    max_order_ = 0;
    for (const Polynomial & p : polynomials_) max_order_ = std::max(max_order_, p.order());
    orders_.assign(max_order_ + 1, {});
    for (const Polynomial & p : polynomials_) orders_[p.order()].push_back(&p);
    engine_.reset();
}
PolynomialSet::PolynomialSet() {}
PolynomialSet::PolynomialSet(const std::vector<Polynomial> & _polynomials, const size_t & _dimension)
: polynomials_(_polynomials), dimension_(_dimension) { construct_orders_(); }
// all terms up to `_order`: orders ascending; inside an order the terms are the non-increasing index tuples,
// enumerated with the first index running fastest (PolynomialSet.cpp:60-111)
PolynomialSet::PolynomialSet(const size_t & _dimension, const size_t & _order) : dimension_(_dimension) {
    polynomials_.push_back(Polynomial(std::vector<size_t>()));
    for (size_t k = 1; k <= _order && dimension_ > 0; k++) {
        std::vector<size_t> c(k, 0);
        while (true) {
            polynomials_.push_back(Polynomial(c));
            // next non-increasing tuple: bump the first position that can still grow, reset the ones before it
            size_t i = 0;
            while (i < k && c[i] + 1 >= dimension_) i++;
            if (i == k) break;
            c[i]++;
            for (size_t j = 0; j < i; j++) c[j] = c[i];
        }
    }
    construct_orders_();
}
PolynomialSet::~PolynomialSet() {}
const std::vector<Polynomial> & PolynomialSet::polynomials() const { return polynomials_; }
const size_t & PolynomialSet::dimension() const { return dimension_; }
const size_t & PolynomialSet::max_order() const { return max_order_; }
const std::vector<std::vector<const Polynomial *>> & PolynomialSet::orders() const { return orders_; }
const Polynomial & PolynomialSet::operator[](const size_t & index) const { return polynomials_[index]; }
std::vector<at::Tensor> PolynomialSet::views(const at::Tensor & x) const {
    std::vector<at::Tensor> out(max_order_ + 1);
    int64_t start = 0;
    for (size_t i = 0; i <= max_order_; i++) { out[i] = x.slice(-1, start, start + (int64_t)orders_[i].size()); start += (int64_t)orders_[i].size(); }
    return out;
}
const SAPSet & PolynomialSet::engine() const {
    if (!engine_) {
        std::vector<SAP> saps;
        for (const Polynomial & p : polynomials_) {
            std::vector<std::pair<size_t, size_t>> coords;
            for (size_t c : p.coords()) coords.emplace_back(0, c);
            saps.push_back(SAP(coords));
        }
        engine_ = std::make_shared<SAPSet>(saps, 0, std::vector<size_t>{dimension_});
    }
    return *engine_;
}
void PolynomialSet::check_(const at::Tensor & x, const char * who_) const {
    const std::string who = std::string("tchem::polynomial::PolynomialSet::") + who_;
    if (x.dim() != 1 && x.dim() != 2) throw std::invalid_argument(who + ": x must be a vector");
    if (x.size(-1) != (int64_t)dimension_) throw std::invalid_argument(who + ": x must have a same dimension as the coordinates");
}
at::Tensor PolynomialSet::operator()(const at::Tensor & x) const { check_(x, "operator()"); return engine()({x}); }
at::Tensor PolynomialSet::Jacobian_(const at::Tensor & x) const {
    check_(x, "Jacobian_");
    at::Tensor J;
    engine().Jacobian_({x}, J);
    return J;
}
at::Tensor PolynomialSet::Jacobian(const at::Tensor & x) const { return Jacobian_(x); }
at::Tensor PolynomialSet::Jacobian2nd_(const at::Tensor & x) const {
    check_(x, "Jacobian2nd_");
    at::Tensor K;
    engine().Jacobian2nd_({x}, K);
    return K;
}
at::Tensor PolynomialSet::Jacobian2nd(const at::Tensor & x) const { return Jacobian2nd_(x); }

} } // namespace tchem::polynomial

// ---- tchem::utility (source/utility.cpp) -----------------------------------------------------------------------
namespace tchem { namespace utility {
std::vector<double> tensor2vector(const at::Tensor & tensor) {
    at::Tensor t = tensor.detach().to(at::kCPU, at::kDouble).contiguous().view({-1});
    return std::vector<double>(t.data_ptr<double>(), t.data_ptr<double>() + t.numel());
}
CL::utility::matrix<double> tensor2matrix(const at::Tensor & tensor) {
    if (tensor.dim() != 2) throw std::invalid_argument("tchem::utility::tensor2matrix: tensor must be a matrix");
    at::Tensor t = tensor.detach().to(at::kCPU, at::kDouble).contiguous();
    CL::utility::matrix<double> m(t.size(0), t.size(1));
    for (int64_t i = 0; i < t.size(0); i++)
    for (int64_t j = 0; j < t.size(1); j++) m[i][j] = t.data_ptr<double>()[i * t.size(1) + j];
    return m;
}
at::Tensor read_vector(std::ifstream & ifs) {
    std::vector<double> v;
    std::string tok;
    while (ifs >> tok) {
        try { v.push_back(std::stod(tok)); } catch (...) { break; }
    }
    return at::tensor(v, at::TensorOptions().dtype(at::kDouble));
}
at::Tensor read_vector(const std::string & file) {
    std::ifstream ifs(file);
    if (!ifs.good()) throw CL::utility::file_error(file);
    return read_vector(ifs);
}
size_t NParameters(const std::vector<at::Tensor> & parameters) {
    size_t n = 0;
    for (const at::Tensor & p : parameters) n += (size_t)p.numel();
    return n;
}
double NetGradNorm(const std::vector<at::Tensor> & parameters) {
    double norm = 0.0;
    for (const at::Tensor & p : parameters) if (p.grad().defined()) norm += p.grad().abs().sum().item<double>();
    return norm;
}
} }


// ---- normal modes (reference source/chem/normal_mode.cpp) ------------------------------
// Batched device operations: cuBLAS batched products for the reductions, cuSOLVER (through ATen) for the Cholesky
// factorisations and the symmetric eigen-solves. IntNormalMode::kernel performs the reduction LAPACK dsygv (itype 3)
// performs behind the reference's Fortran shim: G = L L^T, C = L^T H L, C y = lambda y, x = L y.
namespace tchem { namespace chem {

namespace {
at::Tensor to_device_batched(const at::Tensor & t, bool batched) {
    if (tchem_device_count() <= 0) throw std::runtime_error("tchem::chem: no CUDA device (this library has no CPU path)");
    at::Tensor b = (batched ? t : t.unsqueeze(0)).to(at::kDouble);
    return (b.is_cuda() ? b : b.to(at::Device(at::kCUDA, c10::cuda::current_device()))).clone();
}
at::Tensor signed_sqrt(const at::Tensor & x) { return at::where(x > 0, x.abs().sqrt(), -x.abs().sqrt()); }
at::Tensor mass_vector(const std::vector<double> & masses, const at::Tensor & like) {
    return at::tensor(masses, at::TensorOptions().dtype(at::kDouble)).to(like.device()).repeat_interleave(3);
}
}

CartNormalMode::CartNormalMode() {}
CartNormalMode::CartNormalMode(const std::vector<double> & _masses, const at::Tensor & _Hessian) {
    if (_Hessian.dim() != 2 && _Hessian.dim() != 3) throw std::invalid_argument("tchem::chem::CartNormalMode: Hessian must be a matrix");
    if (_Hessian.size(-1) != _Hessian.size(-2)) throw std::invalid_argument("tchem::chem::CartNormalMode: Hessian must be a square matrix");
    if ((int64_t)(3 * _masses.size()) != _Hessian.size(-1)) throw std::invalid_argument("tchem::chem::CartNormalMode: inconsistent dimension between masses and Hessian");
    masses_ = _masses;
    host_ = !_Hessian.is_cuda();
    batched_ = _Hessian.dim() == 3;
    Hessian_ = to_device_batched(_Hessian, batched_);
}
CartNormalMode::~CartNormalMode() {}
at::Tensor CartNormalMode::out(const at::Tensor & t) const {
    at::Tensor r = batched_ ? t : t[0];
    return host_ ? r.cpu() : r;
}
void CartNormalMode::kernel() {
    at::Tensor s = mass_vector(masses_, Hessian_).sqrt();
    at::Tensor H = Hessian_ / s.unsqueeze(0).unsqueeze(2) / s.unsqueeze(0).unsqueeze(1);
    at::Tensor w, v;
    std::tie(w, v) = at::linalg_eigh(H, "L");
    at::Tensor modes = v.transpose(1, 2);
    // rule out the 6 modes of smallest |eigenvalue| (translations and rotations), then ascending frequency
    at::Tensor keep = at::argsort(w.abs(), /*stable=*/true, 1, false).slice(1, 6);
    at::Tensor freq = signed_sqrt(w.gather(1, keep));
    modes = modes.gather(1, keep.unsqueeze(2).expand({-1, -1, modes.size(2)})) / s.unsqueeze(0).unsqueeze(1);
    at::Tensor order = at::argsort(freq, /*stable=*/true, 1, false);
    frequency_ = freq.gather(1, order);
    cartmode_ = modes.gather(1, order.unsqueeze(2).expand({-1, -1, modes.size(2)}));
    ready_ = true;
}
at::Tensor CartNormalMode::frequency() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::CartNormalMode::frequency");
    return out(frequency_);
}
at::Tensor CartNormalMode::cartmode() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::CartNormalMode::cartmode");
    return out(cartmode_);
}

IntNormalMode::IntNormalMode() {}
IntNormalMode::IntNormalMode(const std::vector<double> & _masses, const at::Tensor & _Jacobian, const at::Tensor & _Hessian) {
    if (_Jacobian.dim() != 2 && _Jacobian.dim() != 3) throw std::invalid_argument("tchem::chem::IntNormalMode: Jacobian must be a matrix");
    if ((int64_t)(3 * _masses.size()) != _Jacobian.size(-1)) throw std::invalid_argument("tchem::chem::IntNormalMode: inconsistent dimension between masses and Jacobian");
    if (_Hessian.dim() != _Jacobian.dim()) throw std::invalid_argument("tchem::chem::IntNormalMode: Hessian must be a matrix");
    if (_Hessian.size(-1) != _Hessian.size(-2)) throw std::invalid_argument("tchem::chem::IntNormalMode: Hessian must be a square matrix");
    if (_Jacobian.size(-2) != _Hessian.size(-1)) throw std::invalid_argument("tchem::chem::IntNormalMode: inconsistent dimension between Jacobian and Hessian");
    masses_ = _masses;
    host_ = !_Jacobian.is_cuda();
    batched_ = _Jacobian.dim() == 3;
    Jacobian_ = to_device_batched(_Jacobian, batched_);
    Hessian_ = to_device_batched(_Hessian, batched_).to(Jacobian_.device());
}
IntNormalMode::~IntNormalMode() {}
void IntNormalMode::kernel() {
    at::Tensor minv = 1.0 / mass_vector(masses_, Jacobian_);
    at::Tensor G = at::matmul(Jacobian_ * minv.unsqueeze(0).unsqueeze(1), Jacobian_.transpose(1, 2));
    at::Tensor L = at::linalg_cholesky(G);
    at::Tensor C = at::matmul(L.transpose(1, 2), at::matmul(Hessian_, L));
    C = 0.5 * (C + C.transpose(1, 2));
    at::Tensor w, Y;
    std::tie(w, Y) = at::linalg_eigh(C, "L");
    frequency_ = signed_sqrt(w);
    intmode_ = at::matmul(L, Y).transpose(1, 2).contiguous();
    Linv_ = at::cholesky_solve(intmode_.transpose(1, 2), L).transpose(1, 2).contiguous();
    at::Tensor Lj = at::linalg_cholesky(at::matmul(Jacobian_, Jacobian_.transpose(1, 2)));
    cartmode_ = at::matmul(at::cholesky_solve(intmode_.transpose(1, 2), Lj).transpose(1, 2), Jacobian_);
    ready_ = true;
}
at::Tensor IntNormalMode::intmode() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::IntNormalMode::intmode");
    return out(intmode_);
}
at::Tensor IntNormalMode::Linv() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::IntNormalMode::Linv");
    return out(Linv_);
}

SANormalMode::SANormalMode() {}
SANormalMode::SANormalMode(const std::vector<double> & _masses, const std::vector<at::Tensor> & _Jacobians, const std::vector<at::Tensor> & _Hessians) {
    if (_Jacobians.size() != _Hessians.size()) throw std::invalid_argument("tchem::chem::SANormalMode: inconsistent number of irreducibles between Jacobian and Hessian");
    for (size_t i = 0; i < _Jacobians.size(); i++) parts_.emplace_back(_masses, _Jacobians[i], _Hessians[i]);
}
SANormalMode::~SANormalMode() {}
void SANormalMode::kernel() { for (auto & p : parts_) p.kernel(); ready_ = true; }
std::vector<at::Tensor> SANormalMode::frequencies() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::SANormalMode::frequencies");
    std::vector<at::Tensor> v; for (const auto & p : parts_) v.push_back(p.frequency()); return v;
}
std::vector<at::Tensor> SANormalMode::intmodes() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::SANormalMode::intmodes");
    std::vector<at::Tensor> v; for (const auto & p : parts_) v.push_back(p.intmode()); return v;
}
std::vector<at::Tensor> SANormalMode::Linvs() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::SANormalMode::Linvs");
    std::vector<at::Tensor> v; for (const auto & p : parts_) v.push_back(p.Linv()); return v;
}
std::vector<at::Tensor> SANormalMode::cartmodes() const {
    if (!ready_) throw CL::utility::not_ready("tchem::chem::SANormalMode::cartmodes");
    std::vector<at::Tensor> v; for (const auto & p : parts_) v.push_back(p.cartmode()); return v;
}
size_t SANormalMode::NIrreds() const { return parts_.size(); }

} }
