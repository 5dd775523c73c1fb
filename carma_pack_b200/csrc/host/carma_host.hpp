// carma_host.hpp -- host C++ classes with the reference's names and method signatures, whose bodies
// call the GPU through the C ABI of include/carma_b200.h.  This is what a C++ user of carma_pack
// links against instead of kfilter.cpp / carpack.cpp / carmcmc.cpp, and what the `_carmcmc` Python
// module (pymodule.cpp) re-exports.
//
// Reference interfaces mirrored (paths relative to /root/reference/src):
//   KalmanFilter<>, KalmanFilter1, KalmanFilterp     include/kfilter.hpp:27-389
//   CARMA_Base<>, CAR1, CARp, ZCAR, CARMA, ZCARMA     include/carpack.hpp:51-461
//   RunCar1Sampler, RunCarmaSampler                  include/carmcmc.hpp, carmcmc.cpp:30-177
// Armadillo is not available in this image, so arma::vec / arma::cx_vec in the reference signatures
// are std::vector<double> / std::vector<std::complex<double>> here (the Boost.Python surface of the
// reference already uses exactly these std types: boost_python_wrapper.cpp:32-43).
#pragma once
#include <complex>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "carma_b200.h"
#include "carma_steps.hpp"

namespace carma_host {

typedef std::vector<std::vector<double> > vecvecD;
typedef std::vector<std::complex<double> > vecC;

// throws std::runtime_error with carma_last_error() when rc != CARMA_OK
void check(int rc, const char* what);

// Seed of the Philox streams used by the samplers (the reference seeds a global mt19937 with
// time(NULL), random.cpp:20).  Every RunCar*Sampler call consumes one "run index".
void set_seed(uint64_t seed);
uint64_t next_run_seed();

// RAII owner of a carma_series_t (sorted, de-duplicated copy of the data: kfilter.hpp:43-76)
class DeviceSeries {
public:
    DeviceSeries(const vecD& time, const vecD& y, const vecD& yerr, int device = 0);
    ~DeviceSeries();
    DeviceSeries(const DeviceSeries&) = delete;
    DeviceSeries& operator=(const DeviceSeries&) = delete;
    carma_series_t handle() const { return h_; }
    const vecD& time() const { return time_; }
    const vecD& y() const { return y_; }
    const vecD& yerr() const { return yerr_; }
    size_t size() const { return time_.size(); }

private:
    carma_series_t h_ = nullptr;
    vecD time_, y_, yerr_;
};

// ------------------------------------------------------------------------------------------------
// Kalman filters (kfilter.hpp)
// ------------------------------------------------------------------------------------------------
template <class OmegaType>
class KalmanFilter {
public:
    vecD mean;  // kfilter.hpp:31-32
    vecD var;

    KalmanFilter(const vecD& time, const vecD& y, const vecD& yerr);
    virtual ~KalmanFilter() {}

    void SetSigsqr(double sigsqr) { sigsqr_ = sigsqr; }
    double GetSigsqr() const { return sigsqr_; }
    virtual void SetOmega(OmegaType omega) { omega_ = omega; }
    OmegaType GetOmega() const { return omega_; }
    vecD GetTime() const { return series_->time(); }
    vecD GetTimeSeries() const { return series_->y(); }
    vecD GetTimeSeriesErr() const { return series_->yerr(); }
    vecD GetMeanSvec() const { return mean; }
    vecD GetVarSvec() const { return var; }

    // kfilter.hpp:126-132: Reset() + (ny-1) x Update(), one GPU call
    void Filter();
    // kfilter.cpp:72-135 / 218-286
    std::pair<double, double> Predict(double time);
    // batched form of Predict: one GPU thread per requested time
    void PredictMany(const vecD& times, vecD& pmean, vecD& pvar);
    // kfilter.hpp:135-184: conditional simulation (sequential insert-and-predict)
    vecD Simulate(vecD time);

protected:
    virtual void params(double& sigsqr, vecC& omega, vecD& ma) const = 0;
    std::shared_ptr<DeviceSeries> series_;
    double sigsqr_ = 1.0;
    OmegaType omega_;
};

class KalmanFilter1 : public KalmanFilter<double> {
public:
    KalmanFilter1(const vecD& time, const vecD& y, const vecD& yerr) : KalmanFilter<double>(time, y, yerr) { omega_ = 1.0; }
    KalmanFilter1(const vecD& time, const vecD& y, const vecD& yerr, double sigsqr, double omega)
        : KalmanFilter<double>(time, y, yerr) { sigsqr_ = sigsqr; omega_ = omega; }

protected:
    void params(double& sigsqr, vecC& omega, vecD& ma) const override;
};

class KalmanFilterp : public KalmanFilter<vecC> {
public:
    KalmanFilterp(const vecD& time, const vecD& y, const vecD& yerr) : KalmanFilter<vecC>(time, y, yerr) {}
    KalmanFilterp(const vecD& time, const vecD& y, const vecD& yerr, double sigsqr, const vecC& omega, const vecD& ma_coefs)
        : KalmanFilter<vecC>(time, y, yerr) { sigsqr_ = sigsqr; omega_ = omega; SetMA(ma_coefs); }
    void SetMA(vecD ma_coefs) { ma_coefs_ = ma_coefs; }
    vecD GetMA() const { return ma_coefs_; }

protected:
    void params(double& sigsqr, vecC& omega, vecD& ma) const override;
    vecD ma_coefs_;
};

// ------------------------------------------------------------------------------------------------
// Model parameter classes (carpack.hpp)
// ------------------------------------------------------------------------------------------------
// CARMA_Base<> of carpack.hpp:51-248: a Parameter<arma::vec> (parameters.hpp:96-169), so it plugs into AdaptiveMetro /
// ExchangeStep / Sampler (carma_steps.hpp) exactly like the reference's classes.
class CARMA_Base : public Parameter<vecD> {
public:
    CARMA_Base(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int kind, int p, int q,
               double temperature = 1.0);
    virtual ~CARMA_Base() {}

    int Dimension() const;
    // carpack.hpp:118-126 / 444-456
    double LogPrior(const vecD& theta) const;
    // carpack.hpp:131-176: one LogDensity on the GPU (the value is remembered for Save(), see below)
    double LogDensity(vecD theta) override;
    // batched LogDensity: rows of theta, one kernel launch
    vecD LogDensityBatch(const vecvecD& theta) const;
    // carpack.hpp:178-191 / carpack.cpp:116-130 / 314-374, evaluated through LogDensity == -inf
    bool CheckPriorBounds(const vecD& theta);
    // carpack.cpp:38-83, 175-230, 416-477, 586-644: drawn on the device until the log-density is finite; the
    // log-density of the returned value is cached, so the Save() that follows costs no second filter run
    vecD StartingValue() override;
    // carpack.cpp:233-265, 479-512, 646-678
    vecD SetStartingValue(vecD init) override;
    // carpack.hpp:90-108: stores the value and its log-posterior.  The reference recomputes the latter from the Kalman
    // filter left behind by the LogDensity(new_value) call that preceded it; here that call's result is cached.
    void Save(vecD new_value) override;
    // value + cached log-posterior in one go (used by ExchangeStep; the reference does Save + SetLogDensity)
    void SetValue(const vecD& value, double logpost) { value_ = value; log_posterior_ = logpost; }
    std::string StringValue() override;
    // carpack.hpp:201-207
    void SetPrior(double max_stdev) { prior_.max_stdev = max_stdev; }
    void SetKappaBounds(double lo, double hi) { prior_.kappa_low = lo; prior_.kappa_high = hi; }
    void SetMLE(bool ignore_prior) { ignore_prior_ = ignore_prior; }
    vecD GetTime() const { return series_->time(); }
    vecD GetTimeSeries() const { return series_->y(); }
    vecD GetTimeSeriesErr() const { return series_->yerr(); }
    const carma_prior_t& prior() const { return prior_; }
    std::shared_ptr<DeviceSeries> series() const { return series_; }
    int kind() const { return kind_; }
    int p() const { return p_; }
    int q() const { return q_; }

    // boost_python_wrapper.cpp:49-73 surface
    double getLogPrior(vecD theta) const { return LogPrior(theta); }
    double getLogDensity(vecD theta) { return LogDensity(theta); }
    vecvecD getSamples() const { return samples_; }

    // filled by the samplers (parameters.hpp:136-147)
    void SetSamples(vecvecD samples, vecD logposts) { samples_ = std::move(samples); logposts_ = std::move(logposts); }
    // diagnostics of the run that produced the samples (steps.hpp:255-264, 365-370)
    vecD accept_rates, exchange_rates;
    // Philox chain index used by StartingValue() (one per chain of a host-built ensemble)
    void SetChainIndex(uint32_t chain) { chain_ = chain; }

protected:
    int kind_, p_, q_;
    std::shared_ptr<DeviceSeries> series_;
    carma_prior_t prior_;
    bool ignore_prior_ = false;
    uint32_t chain_ = 0;
    uint64_t start_draws_ = 0;
    vecD last_theta_;        // argument of the most recent LogDensity call ...
    double last_logpost_ = 0.0;  // ... and its value
};

class CAR1 : public CARMA_Base {
public:
    CAR1(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, double temperature = 1.0)
        : CARMA_Base(track, name, time, y, yerr, CARMA_KIND_CAR1, 1, 0, temperature) {}
};

class CARp : public CARMA_Base {
public:
    CARp(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int p, double temperature = 1.0)
        : CARMA_Base(track, name, time, y, yerr, CARMA_KIND_CARP, p, 0, temperature) {}
    // carpack.cpp:137-172 and 377-409 (host-side scalar helpers; the GPU path has its own copy)
    vecC ARRoots(const vecD& theta) const;
    vecC ExtractAR(const vecD& theta) const { return ARRoots(theta); }  // carpack.hpp:310-313
    double Variance(const vecC& alpha_roots, const vecD& ma_coefs, double sigma, double dt = 0.0) const;
    // carpack.hpp:314 (CARp: [1, 0, ...]); CARMA: carpack.cpp:522-580; ZCARMA: carpack.cpp:687-698
    virtual vecD ExtractMA(const vecD& theta) const;
    // carpack.hpp:316-319, 391-395
    double ExtractSigsqr(const vecD& theta) const { return theta[0] * theta[0] / Variance(ARRoots(theta), ExtractMA(theta), 1.0); }

protected:
    CARp(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int kind, int p, int q,
         double temperature)
        : CARMA_Base(track, name, time, y, yerr, kind, p, q, temperature) {}
};

class ZCAR : public CARp {  // evaluated as CAR(p) by the reference (carpack.hpp:314, 335, 362)
public:
    ZCAR(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int p, double temperature = 1.0)
        : CARp(track, name, time, y, yerr, CARMA_KIND_ZCAR, p, 0, temperature) {}
};

class CARMA : public CARp {
public:
    CARMA(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int p, int q,
          double temperature = 1.0);
    vecD ExtractMA(const vecD& theta) const override;
};

class ZCARMA : public CARp {
public:
    ZCARMA(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int p, double temperature = 1.0)
        : CARp(track, name, time, y, yerr, CARMA_KIND_ZCARMA, p, 0, temperature) {}
    vecD ExtractMA(const vecD& theta) const override;
};

// ------------------------------------------------------------------------------------------------
// Samplers (carmcmc.hpp)
// ------------------------------------------------------------------------------------------------
std::shared_ptr<CAR1> RunCar1Sampler(int sample_size, int burnin, vecD time, vecD y, vecD yerr, int thin = 1,
                                     const vecD& init = vecD());
std::shared_ptr<CARp> RunCarmaSampler(int sample_size, int burnin, vecD time, vecD y, vecD yerr, int p, int q,
                                      int nwalkers, bool do_zcarma = false, int thin = 1, const vecD& init = vecD());
// Many independent ensembles in one launch (the reference runs one): returns one object per ensemble.
std::vector<std::shared_ptr<CARp> > RunCarmaSamplerEnsembles(int n_ensembles, int sample_size, int burnin, vecD time,
                                                             vecD y, vecD yerr, int p, int q, int nwalkers,
                                                             bool do_zcarma = false, int thin = 1,
                                                             const vecD& init = vecD());

}  // namespace carma_host
