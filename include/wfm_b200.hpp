// wfm_b200.hpp -- header-only C++ host mirror of the reference's Java classes over the C ABI.
//
// The reference is compiled (Java) code and this image has no JDK, so the class surface a microTiPi
// user sees is restated in C++ above include/wfm_b200.h with the reference's names, argument
// meaning and error behaviour:
//   microtipi::MicroscopeModel  <- microscopy/MicroscopeModel.java:40-107
//   microtipi::WideFieldModel   <- epifluorescence/WideFieldModel.java (WFM)
// IllegalArgumentException -> std::invalid_argument; any other failure -> std::runtime_error
// (never swallowed, quirk Q7).  Only the TiPi members the callers use are modelled
// (DoubleShapedVectorSpace identity dispatch of WFM:399-422, PSF_Estimation.java:117,202-217).
#pragma once
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "wfm_b200.h"

namespace microtipi {

class DoubleShapedVectorSpace {
public:
    explicit DoubleShapedVectorSpace(int n) : n_(n) {}
    int getNumber() const { return n_; }
private:
    int n_;
};

class DoubleShapedVector {
public:
    DoubleShapedVector(const DoubleShapedVectorSpace* owner, std::vector<double> data) : owner_(owner), data_(std::move(data)) {}
    const DoubleShapedVectorSpace* getOwner() const { return owner_; }
    const DoubleShapedVectorSpace* getSpace() const { return owner_; }
    bool belongsTo(const DoubleShapedVectorSpace* s) const { return s == owner_; }
    int getNumber() const { return (int)data_.size(); }
    double get(int i) const { return data_[i]; }
    void set(int i, double v) { data_[i] = v; }
    std::vector<double>& getData() { return data_; }
    const std::vector<double>& getData() const { return data_; }
    double norm2() const { double s = 0; for (double v : data_) s += v * v; return std::sqrt(s); }
private:
    const DoubleShapedVectorSpace* owner_;
    std::vector<double> data_;
};

// MicroscopeModel.java:40-107
class MicroscopeModel {
public:
    virtual ~MicroscopeModel() = default;
    virtual void computePsf() = 0;                                                        // :103
    virtual DoubleShapedVector apply_Jacobian(const void* grad, const DoubleShapedVectorSpace* xspace) = 0;   // :90
    virtual const int* getParametersFlags() const = 0;                                    // :96
    bool isSingle() const { return single_; }
protected:
    MicroscopeModel(int nx, int ny, int nz, double dxy, double dz, bool single)
        : Nx(nx), Ny(ny), Nz(nz), dxy(dxy), dz(dz), single_(single) {}
    int PState = 0;                                                                       // :42
    int Nx, Ny, Nz;
    double dxy, dz;
    bool single_;
    std::unique_ptr<DoubleShapedVectorSpace> parameterSpace[3];                           // :53
    std::unique_ptr<DoubleShapedVector> parameterCoefs[3];                                // :54
};

class WideFieldModel : public MicroscopeModel {
public:
    static constexpr int DEFOCUS = 0, PHASE = 1, MODULUS = 2;                             // WFM:113-121

    // WFM:154-188.  z0/nz_local select a z-slab of the global stack; defaults = whole stack.
    WideFieldModel(int nx, int ny, int nz, int nPhase, int nModulus, double NA, double lambda, double ni, double dxy,
                   double dz, bool radial, bool single, int device = 0, int z0 = 0, int nz_local = -1)
        : MicroscopeModel(nx, ny, nz, dxy, dz, single), NA_(NA), lambda_(lambda), ni_(ni), radial_(radial) {
        if (nz_local < 0) nz_local = nz - z0;
        nzl_ = nz_local;
        const int rc = wfm_create_slab(&h_, nx, ny, nz, z0, nz_local, dxy, dz, single ? WFM_F32 : WFM_F64, device);
        if (rc != WFM_OK) raise(rc, wfm_last_error(nullptr));                             // WFM:158-160
        lambda_ni_ = ni / lambda;                                                         // WFM:166
        check(wfm_set_optics(h_, NA, lambda, ni));                                        // computeMaskPupil() WFM:174
        nModulus_ = nModulus < 1 ? 1 : nModulus;                                          // WFM:176-179
        nPhase_ = nPhase;
        setNModulus();                                                                    // WFM:185
        setNPhase();                                                                      // WFM:186
        setDefocus(std::vector<double>{ni / lambda, deltaX_, deltaY_});                   // WFM:187
    }
    ~WideFieldModel() override { if (h_) wfm_destroy(h_); }
    WideFieldModel(const WideFieldModel&) = delete;
    WideFieldModel& operator=(const WideFieldModel&) = delete;

    wfm_model* handle() const { return h_; }

    void computePsf() override {                                                          // WFM:206-396
        if (PState > 0) return;
        check(wfm_compute_psf(h_));
        PState = 1;
    }
    // getPsf() WFM:1798: host copy (double or float according to `single`), flat order ix + Nx*(iy + Ny*izl)
    void getPsf(void* out) { if (PState < 1) computePsf(); check(wfm_get_psf(h_, out)); }
    void get_cpxPsf(void* out) { if (PState < 1) computePsf(); check(wfm_get_cpx_psf(h_, out)); }   // WFM:1856
    // getMtf() WFM:1807-1828 as intended (the reference loop never terminates): 3-D DFT of the PSF, (2,Nx,Ny,Nz) doubles
    void getMtf(double* out) { check(wfm_get_mtf(h_, out)); PState = 1; }
    // ArrayUtils.roll(getPsf()) of BlindDeconvJob.java:100: the PSF centred in the volume
    void getPsfRolled(void* out) { check(wfm_get_psf_rolled(h_, out)); PState = 1; }
    void markPsfValid() { PState = wfm_psf_state(h_); }

    DoubleShapedVector apply_Jacobian(const void* grad, const DoubleShapedVectorSpace* xspace) override {   // WFM:399-409
        if (xspace && xspace == parameterSpace[DEFOCUS].get()) return apply_J_defocus(grad);
        if (xspace && xspace == parameterSpace[PHASE].get()) return apply_J_phase(grad);
        if (xspace && xspace == parameterSpace[MODULUS].get()) return apply_J_modulus(grad);
        throw std::invalid_argument("DoubleShapedVector grad does not belong to any space");
    }
    void setParam(const DoubleShapedVector& p) {                                          // WFM:412-422
        if (p.getOwner() == parameterSpace[DEFOCUS].get()) setDefocus(p.getData());
        else if (p.getOwner() == parameterSpace[PHASE].get()) setPhaseCoefs(p.getData());
        else if (p.getOwner() == parameterSpace[MODULUS].get()) setModulusCoefs(p.getData());
        else throw std::invalid_argument("DoubleShapedVector param does not belong to any space");
    }
    DoubleShapedVector apply_J_modulus(const void* q) { return apply(WFM_MODULUS, q, MODULUS); }    // WFM:429
    DoubleShapedVector apply_J_phase(const void* q) {                                               // WFM:738
        if (!parameterSpace[PHASE]) throw std::invalid_argument("phase space is empty");
        return apply(WFM_PHASE, q, PHASE);
    }
    DoubleShapedVector apply_J_defocus(const void* q) { return apply(WFM_DEFOCUS, q, DEFOCUS); }    // WFM:1029

    void setDefocus(const std::vector<double>& defoc) {                                   // WFM:1510-1534 / 1543
        if (!parameterSpace[DEFOCUS]) parameterSpace[DEFOCUS].reset(new DoubleShapedVectorSpace(3));
        if ((int)defoc.size() != 1 && (int)defoc.size() != 3) throw std::invalid_argument("bad defocus  parameters");
        if (defoc.size() == 3) { deltaX_ = defoc[1]; deltaY_ = defoc[2]; }
        lambda_ni_ = defoc[0];
        ni_ = lambda_ni_ * lambda_;
        parameterCoefs[DEFOCUS].reset(new DoubleShapedVector(parameterSpace[DEFOCUS].get(), defoc));
        check(wfm_set_defocus(h_, defoc.data(), (int)defoc.size()));
        freeMem();
    }
    void setPupilAxis(double ax, double ay) { setDefocus({ni_ / lambda_, ax, ay}); }      // WFM:1573
    void setNi(double v) { ni_ = v; lambda_ni_ = v / lambda_; setDefocus({ni_ / lambda_, deltaX_, deltaY_}); }   // WFM:1698
    void setModulus(const std::vector<double>& beta) { setNModulus((int)beta.size()); setModulusCoefs(beta); }   // WFM:1616
    void setPhase(const std::vector<double>& alpha) {                                     // WFM:1655-1665
        if (alpha.empty()) {
            nPhase_ = 0; parameterSpace[PHASE].reset(); parameterCoefs[PHASE].reset();
            check(wfm_set_phase(h_, nullptr, 0));                                         // the device side drops its vector too
            freeMem();
            return;
        }
        setNPhase((int)alpha.size());
        setPhaseCoefs(alpha);
    }
    void setNPhase(int n) { nPhase_ = n; setNPhase(); }                                   // WFM:1919
    void setNModulus(int n) { nModulus_ = n; setNModulus(); }                             // WFM:1930
    void setModulusMode(bool reference_last_plane) {                                      // quirk Q1 switch
        check(wfm_set_modulus_mode(h_, reference_last_plane ? WFM_MODULUS_REFERENCE_LAST_PLANE : WFM_MODULUS_INTENDED));
    }

    std::vector<double> getRho() { return pupil(&wfm_get_rho); }                          // WFM:1673
    std::vector<double> getPhi() { return pupil(&wfm_get_phi); }                          // WFM:1713
    std::vector<double> getPsi() { return pupil(&wfm_get_psi); }                          // WFM:1723
    double getLambda() const { return lambda_; }
    double getNi() const { return ni_; }
    std::vector<double> getDefocus() { if (PState < 1) computePsf(); return {lambda_ni_, deltaX_, deltaY_}; }   // WFM:1761
    int getNZern() const { return Nzern_; }
    int getNModulus() const { return parameterCoefs[MODULUS]->getNumber(); }              // WFM:1981
    int getNPhase() const { return parameterCoefs[PHASE] ? parameterCoefs[PHASE]->getNumber() : 0; }   // WFM:1988
    const DoubleShapedVectorSpace* space(int flag) const { return parameterSpace[flag].get(); }
    const DoubleShapedVector* coefs(int flag) const { return parameterCoefs[flag].get(); }
    const int* getParametersFlags() const override { static const int f[3] = {DEFOCUS, PHASE, MODULUS}; return f; }   // WFM:2000
    void freeMem() { PState = 0; wfm_invalidate(h_); }                                    // WFM:1970

private:
    void setNModulus() {                                                                  // WFM:1939-1961
        if (nModulus_ < 1) nModulus_ = 1;
        parameterSpace[MODULUS].reset(new DoubleShapedVectorSpace(nModulus_));
        const int off = radial_ ? 1 : 3;
        Nzern_ = parameterSpace[PHASE] ? std::max(parameterSpace[PHASE]->getNumber() + off, nModulus_) : nModulus_;
        check(wfm_build_basis(h_, Nzern_, radial_ ? 1 : 0));                              // computeZernike() WFM:194
        std::vector<double> beta(nModulus_, 0.0);
        beta[0] = 1.0;
        setModulusCoefs(beta);
    }
    void setNPhase() {                                                                    // WFM:1899-1914
        if (nPhase_ > 0) {
            parameterSpace[PHASE].reset(new DoubleShapedVectorSpace(nPhase_));
            Nzern_ = std::max(nPhase_ + (radial_ ? 1 : 3), parameterSpace[MODULUS]->getNumber());
            check(wfm_build_basis(h_, Nzern_, radial_ ? 1 : 0));
            setPhaseCoefs(std::vector<double>(nPhase_, 0.0));
        } else {
            parameterSpace[PHASE].reset();
            parameterCoefs[PHASE].reset();
        }
    }
    void setPhaseCoefs(const std::vector<double>& a) {                                    // WFM:1625-1649
        if (!parameterSpace[PHASE] || (int)a.size() != parameterSpace[PHASE]->getNumber())
            throw std::invalid_argument("phase parameter does not belong to the right space  ");
        parameterCoefs[PHASE].reset(new DoubleShapedVector(parameterSpace[PHASE].get(), a));
        check(wfm_set_phase(h_, a.data(), (int)a.size()));
        freeMem();
    }
    void setModulusCoefs(const std::vector<double>& b) {                                  // WFM:1588-1610
        if ((int)b.size() != parameterSpace[MODULUS]->getNumber())
            throw std::invalid_argument("DoubleShapedVector beta does not belong to the modulus space");
        parameterCoefs[MODULUS].reset(new DoubleShapedVector(parameterSpace[MODULUS].get(), b));
        check(wfm_set_modulus(h_, b.data(), (int)b.size()));
        freeMem();
    }
    DoubleShapedVector apply(int param, const void* q, int flag) {
        std::vector<double> out(parameterSpace[flag]->getNumber());
        check(wfm_apply_jacobian(h_, param, q, out.data(), (int)out.size()));
        PState = wfm_psf_state(h_);
        return DoubleShapedVector(parameterSpace[flag].get(), std::move(out));
    }
    std::vector<double> pupil(int (*fn)(wfm_model*, double*)) {
        if (PState < 1) computePsf();                                                     // WFM:1674-1676
        std::vector<double> out((size_t)Nx * Ny);
        check(fn(h_, out.data()));
        return out;
    }
    void check(int rc) const { if (rc != WFM_OK) raise(rc, wfm_last_error(h_)); }
    [[noreturn]] static void raise(int rc, const char* msg) {
        if (rc == WFM_ERR_INVALID_ARG || rc == WFM_ERR_UNSUPPORTED) throw std::invalid_argument(msg);
        throw std::runtime_error(std::string("wfm_b200: ") + msg + " (status " + std::to_string(rc) + ")");
    }

    wfm_model* h_ = nullptr;
    double NA_, lambda_, ni_, lambda_ni_ = 0, deltaX_ = 0, deltaY_ = 0;
    bool radial_;
    int nPhase_ = 0, nModulus_ = 1, Nzern_ = 4, nzl_ = 0;
};

// TiPi mitiv.conv.WeightedConvolutionCost as PSF_Estimation drives it (PSF_Estimation.java:147-150,157,206): the
// object is the kernel of the operator, the PSF is the variable (restated semantics, see wfm_conv.cuh).
class WeightedConvolutionCost {
public:
    WeightedConvolutionCost(int nx, int ny, int nz, int device = 0) : vox_((size_t)nx * ny * nz) {     // build(space) :147
        const int rc = wfm_conv_create(&c_, nx, ny, nz, WFM_F64, device);
        if (rc != WFM_OK) raise(rc, wfm_conv_last_error(nullptr));
    }
    ~WeightedConvolutionCost() { if (c_) wfm_conv_destroy(c_); }
    WeightedConvolutionCost(const WeightedConvolutionCost&) = delete;
    WeightedConvolutionCost& operator=(const WeightedConvolutionCost&) = delete;
    void setPSF(const double* obj) { check(wfm_conv_set_object(c_, obj)); }               // setPSF(obj, {0,0,0}) :145,148
    void setData(const double* data) { check(wfm_conv_set_data(c_, data)); }              // :149
    void setWeights(const double* w) { check(wfm_conv_set_weights(c_, w)); }              // :150 (nullptr = unit weights)
    double computeCostAndGradient(double alpha, const double* x, double* gx, bool clr) {  // :157,206
        double cost = 0.0;
        check(wfm_conv_cost_and_gradient(c_, alpha, x, gx, clr ? 1 : 0, &cost));
        return cost;
    }
    // One COMPUTE_FG evaluation of PSF_Estimation.fitPSF (:202-217) on the device; returns the cost, fills g.
    double evalFG(WideFieldModel& pupil, const DoubleShapedVector& x, std::vector<double>& g, double alpha = 1.0) {
        int flag = -1;
        for (int f = 0; f < 3; ++f) if (pupil.space(f) && x.getOwner() == pupil.space(f)) flag = f;
        if (flag < 0) throw std::invalid_argument("DoubleShapedVector param does not belong to any space");
        g.assign(x.getNumber(), 0.0);
        double cost = 0.0;
        pupil.setParam(x);        // WFM:412-422: parameterCoefs (and ni / deltaX / deltaY for the defocus group) follow x
        const int rc = wfm_eval_fg(pupil.handle(), c_, flag, nullptr, x.getNumber(), alpha, &cost, g.data());
        if (rc != WFM_OK) raise(rc, wfm_last_error(pupil.handle()));
        pupil.markPsfValid();
        return cost;
    }
    size_t voxels() const { return vox_; }
private:
    void check(int rc) const { if (rc != WFM_OK) raise(rc, wfm_conv_last_error(c_)); }
    [[noreturn]] static void raise(int rc, const char* msg) {
        if (rc == WFM_ERR_INVALID_ARG || rc == WFM_ERR_UNSUPPORTED) throw std::invalid_argument(msg);
        throw std::runtime_error(std::string("wfm_b200: ") + msg + " (status " + std::to_string(rc) + ")");
    }
    wfm_conv* c_ = nullptr;
    size_t vox_;
};

}  // namespace microtipi
