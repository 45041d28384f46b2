// Exercises the C++ host mirror (include/wfm_b200.hpp) through the C ABI, the way PSF_Estimation
// drives the reference (PSF_Estimation.java:202-217): setParam -> computePsf -> getPsf -> apply_Jacobian.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "wfm_b200.hpp"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    using namespace microtipi;
    const int N = 32, Nz = 6;
    try { WideFieldModel bad(32, 64, 4, 10, 1, 1.4, 542e-9, 1.518, 64.5e-9, 160e-9, false, false); CHECK(false); }
    catch (const std::invalid_argument& e) { CHECK(std::string(e.what()).find("Nx should equal Ny") != std::string::npos); }   // WFM:158

    WideFieldModel m(N, N, Nz, 10, 4, 1.4, 542e-9, 1.518, 64.5e-9, 160e-9, false, false);
    CHECK(m.getNZern() == 13 && m.getNPhase() == 10 && m.getNModulus() == 4);
    DoubleShapedVector x(m.space(WideFieldModel::PHASE), std::vector<double>{0.1, -0.2, 0.05, 0, 0.3, -0.1, 0, 0.02, 0, -0.04});
    m.setParam(x);                                                     // PSF_Estimation.java:202
    m.computePsf();                                                    // :204
    std::vector<double> psf((size_t)N * N * Nz), rho;
    m.getPsf(psf.data());
    rho = m.getRho();
    double e = 0, r2 = 0;
    for (double v : psf) e += v;
    for (double v : rho) r2 += v * v;
    CHECK(std::fabs(r2 - 1.0) < 1e-12);                                // orthonormal basis, normalised beta
    CHECK(std::fabs(e - 1.0) < 1e-12);                                 // Parseval with PSFnorm (WFM:284,327)

    std::vector<double> q(psf.size());
    for (size_t i = 0; i < q.size(); ++i) q[i] = std::sin(0.37 * (double)i);
    DoubleShapedVector g = m.apply_Jacobian(q.data(), x.getSpace());   // :217
    CHECK(g.getNumber() == 10 && g.belongsTo(m.space(WideFieldModel::PHASE)));
    // finite-difference check of one component: apply_J_phase is the exact gradient of sum(q*psf)
    auto cost = [&](const std::vector<double>& a) {
        m.setParam(DoubleShapedVector(m.space(WideFieldModel::PHASE), a));
        std::vector<double> p(psf.size());
        m.getPsf(p.data());
        double c = 0;
        for (size_t i = 0; i < p.size(); ++i) c += q[i] * p[i];
        return c;
    };
    std::vector<double> ap = x.getData(), am = x.getData();
    ap[4] += 1e-6; am[4] -= 1e-6;
    const double fd = (cost(ap) - cost(am)) / 2e-6;
    double gmax = 0;
    for (double v : g.getData()) gmax = std::fmax(gmax, std::fabs(v));
    CHECK(std::fabs(fd - g.get(4)) <= 2e-6 * gmax);

    DoubleShapedVectorSpace alien(10);
    try { m.apply_Jacobian(q.data(), &alien); CHECK(false); }
    catch (const std::invalid_argument&) {}                            // WFM:407
    try { m.setDefocus({1.0, 2.0}); CHECK(false); }
    catch (const std::invalid_argument&) {}                            // quirk Q4
    DoubleShapedVector d = m.apply_J_defocus(q.data());
    CHECK(d.getNumber() == 3 && std::isfinite(d.get(0)));
    // rows f1 / f4 through the C++ mirror: the data term and one COMPUTE_FG evaluation (PSF_Estimation.java:147-157,202-217)
    {
        const int Nc = 32, Nzc = 32;
        WideFieldModel pupil(Nc, Nc, Nzc, 10, 1, 1.4, 542e-9, 1.518, 64.5e-9, 160e-9, false, false);
        const size_t vox = (size_t)Nc * Nc * Nzc;
        std::vector<double> obj(vox, 0.0), data(vox), psf0(vox), rolled(vox), mtf(2 * vox), gq(vox);
        obj[0] = 1.0; obj[1] = 0.5; obj[Nc] = 0.25;                    // a small blob at the origin
        pupil.getPsf(psf0.data());
        for (size_t i = 0; i < vox; ++i) data[i] = 0.9 * psf0[i];
        WeightedConvolutionCost fdata(Nc, Nc, Nzc);
        fdata.setPSF(obj.data()); fdata.setData(data.data()); fdata.setWeights(nullptr);
        const double c0 = fdata.computeCostAndGradient(1.0, psf0.data(), gq.data(), true);
        CHECK(c0 > 0 && std::isfinite(c0));
        DoubleShapedVector xa(pupil.space(WideFieldModel::PHASE), std::vector<double>(10, 0.01));
        std::vector<double> gfg;
        const double c1 = fdata.evalFG(pupil, xa, gfg);
        // the same evaluation step by step through the host-buffer calls
        pupil.setParam(xa);
        std::vector<double> p1(vox);
        pupil.getPsf(p1.data());
        const double c2 = fdata.computeCostAndGradient(1.0, p1.data(), gq.data(), true);
        DoubleShapedVector g2 = pupil.apply_Jacobian(gq.data(), xa.getSpace());
        CHECK(std::fabs(c1 - c2) <= 1e-12 * std::fabs(c2));
        double num = 0, den = 0;
        for (int k = 0; k < 10; ++k) { num += (gfg[k] - g2.get(k)) * (gfg[k] - g2.get(k)); den += g2.get(k) * g2.get(k); }
        CHECK(std::sqrt(num) <= 1e-12 * std::sqrt(den));
        pupil.getPsfRolled(rolled.data());
        CHECK(rolled[(Nc / 2) + (size_t)Nc * ((Nc / 2) + (size_t)Nc * (Nzc / 2))] == p1[0]);     // origin moved to the centre
        pupil.getMtf(mtf.data());
        double sum = 0;
        for (double v : p1) sum += v;
        CHECK(std::fabs(mtf[0] - sum) <= 1e-12 && std::fabs(mtf[1]) <= 1e-15);                   // DC term = total energy
    }
    std::printf("OK energy=%.15f fd=%.6e g4=%.6e\n", e, fd, g.get(4));
    return 0;
}
