// C++ check of the drop-in classes on a GPU box, in the style of the reference's own test
// programs (test/intcoord/main.cpp, test/SApolynomial/main.cpp): it prints residuals that should
// be close to 0 - and, unlike the reference, exits non-zero when one is not.
//
// usage: test_intcoord.exe <dir>   with <dir> holding IntCoordDef, intcfl, advanced_IntCoordDef,
//        r_default.txt, r_columbus.txt (Cartesian coordinates in bohr, one number per line),
//        sap1.in, sap2.in   (written by tests/test_cpp_host.py from tests/golden/)
#include <fstream>
#include <iostream>

#include <tchem/tchem.hpp>
#include <tchem/chem/normal_mode.hpp>

namespace {
int failures = 0;
void report(const std::string & what, double value, double bound) {
    const bool ok = value <= bound;
    std::cout << (ok ? "  ok   " : "  FAIL ") << what << ": " << value << " (bound " << bound << ")\n";
    if (!ok) failures++;
}
at::Tensor read_vector(const std::string & file) {
    std::ifstream ifs(file);
    if (!ifs.good()) throw std::runtime_error("cannot open " + file);
    std::vector<double> v;
    double x;
    while (ifs >> x) v.push_back(x);
    return at::tensor(v, at::TensorOptions().dtype(at::kDouble));
}
double maxabs(const at::Tensor & t) { return t.numel() ? t.abs().max().item<double>() : 0.0; }
template <typename E, typename F> bool throws(F && f) {
    try { f(); } catch (const E &) { return true; } catch (...) { return false; }
    return false;
}
}

int main(int argc, char ** argv) {
    const std::string dir = argc > 1 ? std::string(argv[1]) + "/" : "./";
    std::cout << "tchem B200 drop-in: module 'intcoord'. Correct routines print values close to 0\n";
    at::Tensor r_cpu = read_vector(dir + "r_default.txt"), r_col_cpu = read_vector(dir + "r_columbus.txt");
    at::Tensor r = r_cpu.cuda(), r_col = r_col_cpu.cuda();

    tchem::IC::IntCoordSet set("default", dir + "IntCoordDef"), set_col("Columbus7", dir + "intcfl");
    std::cout << "q, J, K (" << set.size() << " internal coordinates, " << r.size(0) / 3 << " atoms)\n";
    at::Tensor q0 = set(r), q1, J1, q2, J2, K2;
    std::tie(q1, J1) = set.compute_IC_J(r);
    std::tie(q2, J2, K2) = set.compute_IC_J_K(r);
    report("q from operator() vs compute_IC_J", maxabs(q0 - q1), 1e-9);
    report("q from compute_IC_J vs compute_IC_J_K", maxabs(q1 - q2), 1e-14);
    report("J from compute_IC_J vs compute_IC_J_K", maxabs(J1 - J2), 1e-14);
    // central differences, batched: all 2 * cartdim displaced geometries in one call
    const int64_t cd = r.size(0);
    const double h = 1e-5;
    at::Tensor disp = at::eye(cd, r.options()) * h;
    at::Tensor rp = r.unsqueeze(0) + disp, rm = r.unsqueeze(0) - disp, qp, Jp, qm, Jm;
    std::tie(qp, Jp) = set.compute_IC_J(rp);
    std::tie(qm, Jm) = set.compute_IC_J(rm);
    at::Tensor J_fd = ((qp - qm) / (2 * h)).transpose(0, 1);              // [intdim][cartdim]
    at::Tensor K_fd = ((Jp - Jm) / (2 * h)).permute({1, 2, 0});            // [intdim][cartdim][cartdim]
    report("finite difference vs analytical J", maxabs(J2 - J_fd), 1e-7);
    report("finite difference vs analytical K", maxabs(K2 - K_fd), 1e-6);
    // the Columbus7 file defines the same coordinates
    at::Tensor qc, Jc, Kc, qd, Jd, Kd;
    std::tie(qc, Jc, Kc) = set_col.compute_IC_J_K(r_col);
    std::tie(qd, Jd, Kd) = set.compute_IC_J_K(r_col);
    report("Columbus7 vs default format: q", maxabs(qc - qd), 1e-13);
    report("Columbus7 vs default format: J", maxabs(Jc - Jd), 1e-13);
    report("Columbus7 vs default format: K", maxabs(Kc - Kd), 1e-13);
    // host tensors in, host tensors out
    at::Tensor qh, Jh;
    std::tie(qh, Jh) = set.compute_IC_J(r_cpu);
    report("host-buffer call equals device call", maxabs(qh - q1.cpu()) + maxabs(Jh - J1.cpu()), 0.0);
    if (qh.is_cuda() || Jh.is_cuda()) { std::cout << "  FAIL results of a host call must live on the host\n"; failures++; }

    std::cout << "all 14 invariant displacement types\n";
    tchem::IC::IntCoordSet adv("whatever", dir + "advanced_IntCoordDef");
    at::Tensor q = adv(r), qJ, J, qK, JK, K;
    std::tie(qJ, J) = adv.compute_IC_J(r);
    std::tie(qK, JK, K) = adv.compute_IC_J_K(r);
    double qerror = std::abs(q[0].item<double>() - 1.0);
    auto acc = [&](const at::Tensor & t) { qerror += std::abs(t.item<double>()); };
    acc(at::cos(q[1]) - q[2]); acc(at::cos(q[3]) - q[4]); acc(at::cos(q[5]) - q[6]);
    acc(at::sin(q[7]) - q[8]); acc(at::sin(q[10]) - q[11]); acc(at::cos(q[7]) - q[9]); acc(at::cos(q[10]) - q[12]);
    acc(at::sin(q[13]) - q[14]); acc(at::sin(q[15]) - q[16]); acc(at::sin(q[17]) - q[18]); acc(at::cos(q[17]) - q[19]);
    acc(at::sin(q[1]) * q[19] - q[20]); acc(at::sin(q[1]) * q[18] - q[21]);
    report("identities between the advanced coordinates", qerror, 1e-12);
    std::tie(qp, Jp) = adv.compute_IC_J(rp);
    std::tie(qm, Jm) = adv.compute_IC_J(rm);
    report("advanced: finite difference vs J", maxabs(J - ((qp - qm) / (2 * h)).transpose(0, 1)), 1e-7);
    report("advanced: finite difference vs K", maxabs(K - ((Jp - Jm) / (2 * h)).permute({1, 2, 0})), 1e-6);
    report("advanced: q from the three paths", maxabs(q - qJ) + maxabs(qJ - qK), 1e-9);
    // single InvDisp / IntCoord evaluate like a one-row set
    at::Tensor qi, Ji;
    std::tie(qi, Ji) = adv[7][0].second.compute_IC_J(r);
    report("InvDisp::compute_IC_J equals the set's row", std::abs((qi - qJ[7]).item<double>()) + maxabs(Ji - J[7]), 1e-15);
    std::tie(qi, Ji) = set[8].compute_IC_J(r);
    report("IntCoord::compute_IC_J equals the set's row", std::abs((qi - q1[8]).item<double>()) + maxabs(Ji - J1[8]), 1e-15);

    std::cout << "gradient and Hessian transformations (round trip of test/intcoord/main.cpp:120-149)\n";
    const int64_t n = set.size();
    at::Tensor intgrad = q1 - q1.mean(), intHess = at::eye(n, r.options());
    at::Tensor cartgrad = set.gradient_int2cart(r, intgrad), cartHess = set.Hessian_int2cart(r, intgrad, intHess);
    at::Tensor intgrad0 = set.gradient_cart2int_matrix(r).mv(cartgrad);
    at::Tensor intgrad1 = set.gradient_cart2int(r, cartgrad), intHess1 = set.Hessian_cart2int(r, cartgrad, cartHess);
    at::Tensor cartgrad1 = set.gradient_int2cart(r, intgrad1), cartHess1 = set.Hessian_int2cart(r, intgrad1, intHess1);
    report("int gradient round trip", (intgrad - intgrad0).norm().item<double>() + (intgrad - intgrad1).norm().item<double>(), 1e-9);
    report("int Hessian round trip", (intHess - intHess1).norm().item<double>(), 1e-8);
    report("cart gradient round trip", (cartgrad - cartgrad1).norm().item<double>(), 1e-9);
    report("cart Hessian round trip", (cartHess - cartHess1).norm().item<double>(), 1e-8);

    std::cout << "definition printing round trip\n";
    for (const std::string fmt : {"default", "Columbus7"}) {
        const std::string file = dir + "printed_" + fmt;
        { std::ofstream ofs(file); set_col.print(ofs, fmt); }
        tchem::IC::IntCoordSet again(fmt, file);
        at::Tensor qa, Ja;
        std::tie(qa, Ja) = again.compute_IC_J(r);
        std::tie(qd, Jd) = set_col.compute_IC_J(r);
        report("print(" + fmt + ") -> read back", maxabs(qa - qd) + maxabs(Ja - Jd), 2e-6);   // 6-digit coefficients
    }

    std::cout << "error behaviour\n";
    report("2-D Hessian argument check", throws<std::invalid_argument>([&] { set.Hessian_int2cart(r, intgrad, intgrad); }) ? 0 : 1, 0);
    report("gradient dimension check", throws<std::invalid_argument>([&] { set.gradient_int2cart(r, cartgrad); }) ? 0 : 1, 0);
    report("missing file", throws<CL::utility::file_error>([&] { tchem::IC::IntCoordSet bad("default", dir + "nope"); }) ? 0 : 1, 0);
    report("unknown type", throws<std::invalid_argument>([&] { tchem::IC::InvDisp bad("wagging", {0, 1}); }) ? 0 : 1, 0);

    std::cout << "module 'SApolynomial' (test/SApolynomial/main.cpp:9-25)\n";
    at::Tensor answer1 = at::tensor({1.0, 2.0, 3.0, 4.0, 6.0, 9.0, 25.0, 35.0, 49.0}, r_cpu.options()),
               answer2 = at::tensor({5.0, 7.0, 10.0, 14.0, 15.0, 21.0}, r_cpu.options());
    tchem::polynomial::SAPSet set1(dir + "sap1.in", 0, {2, 2}), set2(dir + "sap2.in", 1, {2, 2});
    std::vector<at::Tensor> xs = {at::tensor({2.0, 3.0}, r_cpu.options()), at::tensor({5.0, 7.0}, r_cpu.options())};
    report("value of symmetry adapted polynomial set", (set1(xs) - answer1).norm().item<double>() + (set2(xs) - answer2).norm().item<double>(), 0.0);
    at::Tensor cat_J;
    auto Ja = set1.Jacobian(xs), Jb = set1.Jacobian_(xs), Jc2 = set1.Jacobian_(xs, cat_J);
    double jerr = 0.0;
    for (size_t i = 0; i < 2; i++) jerr += maxabs(Ja[i] - Jb[i]) + maxabs(Ja[i] - Jc2[i]);
    jerr += maxabs(cat_J - at::cat(Ja, 1));
    // against per-term autograd-free gradients of SAP
    for (size_t k = 0; k < set1.SAPs().size(); k++) {
        auto g = set1[k].gradient(xs);
        jerr += maxabs(at::cat(g) - cat_J[k]);
        jerr += std::abs((set1[k](xs) - answer1[k]).item<double>());
    }
    report("the three Jacobian variants and SAP::gradient agree", jerr, 1e-13);
    // second-order Jacobian: the three variants agree, K is symmetric and matches central differences of Jacobian_
    at::Tensor cat_K;
    auto Ka = set1.Jacobian2nd(xs), Kb = set1.Jacobian2nd_(xs), Kcat = set1.Jacobian2nd_(xs, cat_K);
    double kerr = maxabs(cat_K - cat_K.transpose(-1, -2));
    kerr += maxabs(Ka[0][1] - Kcat[0][1]) + maxabs(Kb[0][0] - at::triu(Kcat[0][0])) + maxabs(Kb[1][1] - at::triu(Kcat[1][1]));
    report("the three Jacobian2nd variants agree, K symmetric", kerr, 1e-13);
    at::Tensor xcat = at::cat(xs);
    double fderr = 0.0;
    for (int64_t c = 0; c < 4; c++) {
        at::Tensor xp = xcat.clone(), xm = xcat.clone();
        xp[c] += 1e-5; xm[c] -= 1e-5;
        at::Tensor Jp, Jm;
        set1.Jacobian_({xp.slice(0, 0, 2), xp.slice(0, 2, 4)}, Jp);
        set1.Jacobian_({xm.slice(0, 0, 2), xm.slice(0, 2, 4)}, Jm);
        fderr = std::max(fderr, maxabs((Jp - Jm) / 2e-5 - cat_K.select(-1, c)));
    }
    report("Jacobian2nd_ = d Jacobian_ / dx (central differences)", fderr, 1e-6);
    // per-term Hessians (at:: arithmetic) = upper triangle of the set's rows
    double herr = 0.0;
    for (size_t k = 0; k < set1.SAPs().size(); k++) {
        at::Tensor hk;
        set1[k].Hessian_(xs, hk);
        herr = std::max(herr, maxabs(hk - at::triu(cat_K[k])));
        auto hb = set1[k].Hessian(xs);
        herr = std::max(herr, maxabs(hb[0][1] - cat_K[k].slice(0, 0, 2).slice(1, 2, 4)));
    }
    report("SAP::Hessian variants = upper triangle of Jacobian2nd_ rows", herr, 1e-13);
    report("constant term only in the totally symmetric irreducible",
           throws<std::invalid_argument>([&] { tchem::polynomial::SAPSet bad(dir + "sap1.in", 1, {2, 2}); }) ? 0 : 1, 0);

    std::cout << "module 'polynomial' (test/polynomial/main.cpp:7-16)\n";
    at::Tensor panswer = at::tensor({1.0, 2.0, 7.0, 4.0, 14.0, 49.0, 8.0, 28.0, 98.0, 343.0, 16.0, 56.0, 196.0, 686.0, 2401.0}, r_cpu.options());
    tchem::polynomial::PolynomialSet pset(2, 4);
    at::Tensor px = at::tensor({2.0, 7.0}, r_cpu.options());
    report("value of polynomial set", (pset(px) - panswer).norm().item<double>(), 0.0);
    tchem::polynomial::PolynomialSet pset3(3, 2);
    const std::vector<std::vector<size_t>> expect = {{}, {0}, {1}, {2}, {0, 0}, {1, 0}, {2, 0}, {1, 1}, {2, 1}, {2, 2}};
    double order_err = pset3.polynomials().size() == expect.size() ? 0.0 : 1.0;
    for (size_t i = 0; i < expect.size() && order_err == 0.0; i++) if (pset3[i].coords() != expect[i]) order_err = 1.0;
    report("generated terms (3 coordinates, order 2) in the reference's order", order_err, 0.0);
    at::Tensor PJ = pset.Jacobian_(px), PK = pset.Jacobian2nd_(px);
    double pfd = 0.0;
    for (int64_t c = 0; c < 2; c++) {
        at::Tensor xp = px.clone(), xm = px.clone();
        xp[c] += 1e-4; xm[c] -= 1e-4;
        pfd = std::max(pfd, maxabs((pset(xp) - pset(xm)) / 2e-4 - PJ.select(-1, c)) / 2401.0);
        pfd = std::max(pfd, maxabs((pset.Jacobian_(xp) - pset.Jacobian_(xm)) / 2e-4 - PK.select(-1, c)) / 2401.0);
    }
    report("polynomial Jacobian / Jacobian2nd vs central differences", pfd, 1e-6);
    report("Polynomial::operator(), gradient_, Hessian_ = rows of the set",
           std::abs((pset[8](px) - panswer[8]).item<double>()) + maxabs(pset[8].gradient_(px) - PJ[8]) + maxabs(pset[8].Hessian_(px) - PK[8]), 1e-12);
    report("dimension check", throws<std::invalid_argument>([&] { pset(at::ones({3}, r_cpu.options())); }) ? 0 : 1, 0);

    {   // one call with host tensors spread over every visible GPU: bit-identical to the one-device result
        const int64_t nb = 8192 * 2;
        at::Tensor rb = r_cpu.unsqueeze(0).repeat({nb, 1}) + 0.01 * at::randn({nb, r_cpu.size(0)}, r_cpu.options());
        at::Tensor qa, Ja, qb, Jb;
        tchem::b200::set_host_devices(1);
        std::tie(qa, Ja) = set.compute_IC_J(rb);
        tchem::b200::set_host_devices(1 << 20);
        std::tie(qb, Jb) = set.compute_IC_J(rb);
        std::cout << "host tensors over " << tchem::b200::host_devices() << " device(s)\n";
        report("compute_IC_J over all devices vs one device", maxabs(Ja - Jb) + maxabs(qa - qb), 0.0);
        tchem::b200::set_host_devices(-1);
    }
    {   // normal modes (reference test/normal_mode/main.cpp:49-72): Cartesian and internal-coordinate (GF) analyses of the
        // same force field give the same frequencies; here on the aniline fixture with a synthetic positive-definite H_q
        const int64_t n = (int64_t)set.size(), natoms = r_cpu.size(0) / 3;
        at::Tensor A = at::randn({n, n}, r_cpu.options());
        at::Tensor Hq = at::matmul(A, A.transpose(0, 1)) / (double)n + at::eye(n, r_cpu.options());
        std::vector<double> masses(natoms);
        for (int64_t i = 0; i < natoms; i++) masses[i] = i % 3 == 0 ? 12.0 : (i % 3 == 1 ? 1.00782503223 : 14.0030740048);
        at::Tensor qv, Jv;
        std::tie(qv, Jv) = set.compute_IC_J(r);
        at::Tensor Hr = set.Hessian_int2cart(r, at::zeros({n}, r.options()), Hq.cuda());
        tchem::chem::CartNormalMode cartvib(masses, Hr);
        report("observables before kernel() throw not_ready", throws<CL::utility::not_ready>([&] { cartvib.frequency(); }) ? 0 : 1, 0);
        cartvib.kernel();
        tchem::chem::IntNormalMode intvib(masses, Jv, Hq.cuda());
        intvib.kernel();
        report("normal modes: Cartesian vs internal frequencies", (cartvib.frequency() - intvib.frequency()).norm().item<double>(), 1e-9);
        report("normal modes: Linv . intmode^T = 1", maxabs(at::matmul(intvib.Linv(), intvib.intmode().transpose(0, 1)) - at::eye(n, r.options())), 1e-9);
        at::Tensor mw = cartvib.cartmode() * at::tensor(masses, r_cpu.options()).repeat_interleave(3).cuda().unsqueeze(0);
        report("normal modes: Cartesian-analysis modes normalised under the mass metric", maxabs(at::matmul(mw, cartvib.cartmode().transpose(0, 1)) - at::eye(n, r.options())), 1e-8);
        report("normal modes: shapes", (intvib.cartmode().size(0) == n && intvib.cartmode().size(1) == r.size(0)) ? 0 : 1, 0);
    }
    std::cout << (failures ? "FAILED " : "passed ") << "(" << failures << " failures)\n";
    return failures ? 1 : 0;
}
