// carma_host.cpp -- bodies of the reference-named host classes: thin calls into the C ABI.
#include "carma_host.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iostream>
#include <numeric>
#include <random>
#include <sstream>

namespace carma_host {

void check(int rc, const char* what) {
    if (rc != CARMA_OK) {
        const char* msg = carma_last_error();
        throw std::runtime_error(std::string(what) + " failed: " + (msg ? msg : "") + " (code " + std::to_string(rc) + ")");
    }
}

// ---- seeding ---------------------------------------------------------------------------------
static uint64_t g_seed = (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();  // random.cpp:20
static uint64_t g_run = 0;
void set_seed(uint64_t seed) { g_seed = seed; g_run = 0; }
uint64_t next_run_seed() {
    // splitmix64 of (seed, run index): independent Philox keys for successive sampler runs
    uint64_t z = g_seed + 0x9E3779B97F4A7C15ull * (++g_run);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- DeviceSeries ----------------------------------------------------------------------------
DeviceSeries::DeviceSeries(const vecD& time, const vecD& y, const vecD& yerr, int device) {
    if (time.size() != y.size() || time.size() != yerr.size()) throw std::invalid_argument("time, y, yerr differ in length");
    // kfilter.hpp:43-76: sort by time, then drop duplicate times (keep the first of each run)
    std::vector<size_t> idx(time.size());
    std::iota(idx.begin(), idx.end(), 0);
    bool sorted = std::is_sorted(time.begin(), time.end());
    if (!sorted) {
        std::cout << "Time vector is not sorted in increasing order. Sorting the data vectors..." << std::endl;
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return time[a] < time[b]; });
    }
    bool dup = false;
    for (size_t k = 0; k < idx.size(); k++) {
        if (k > 0 && time[idx[k]] == time_.back()) { dup = true; continue; }
        time_.push_back(time[idx[k]]);
        y_.push_back(y[idx[k]]);
        yerr_.push_back(yerr[idx[k]]);
    }
    if (dup) std::cout << "Found duplicate values of time, removing them..." << std::endl;
    check(carma_series_create(time_.data(), y_.data(), yerr_.data(), time_.size(), device, &h_), "carma_series_create");
}

DeviceSeries::~DeviceSeries() {
    if (h_) carma_series_destroy(h_);
}

// ---- KalmanFilter ----------------------------------------------------------------------------
template <class O>
KalmanFilter<O>::KalmanFilter(const vecD& time, const vecD& y, const vecD& yerr)
    : series_(std::make_shared<DeviceSeries>(time, y, yerr)) {
    mean.assign(series_->size(), 0.0);
    var.assign(series_->size(), 0.0);
}

static std::vector<double> flatten(const vecC& omega) {
    std::vector<double> o(2 * omega.size());
    for (size_t i = 0; i < omega.size(); i++) { o[2 * i] = omega[i].real(); o[2 * i + 1] = omega[i].imag(); }
    return o;
}

template <class O>
void KalmanFilter<O>::Filter() {
    double s2; vecC om; vecD ma;
    params(s2, om, ma);
    std::vector<double> o = flatten(om);
    mean.assign(series_->size(), 0.0);
    var.assign(series_->size(), 0.0);
    check(carma_filter(series_->handle(), s2, o.data(), ma.data(), (int)om.size(), 1.0, 0.0, mean.data(), var.data()),
          "carma_filter");
}

template <class O>
void KalmanFilter<O>::PredictMany(const vecD& times, vecD& pmean, vecD& pvar) {
    double s2; vecC om; vecD ma;
    params(s2, om, ma);
    std::vector<double> o = flatten(om);
    pmean.assign(times.size(), 0.0);
    pvar.assign(times.size(), 0.0);
    check(carma_predict(series_->handle(), s2, o.data(), ma.data(), (int)om.size(), 1.0, 0.0, times.data(), times.size(),
                        pmean.data(), pvar.data()), "carma_predict");
}

template <class O>
std::pair<double, double> KalmanFilter<O>::Predict(double time) {
    vecD t(1, time), m, v;
    PredictMany(t, m, v);
    return std::pair<double, double>(m[0], v[0]);
}

template <class O>
vecD KalmanFilter<O>::Simulate(vecD time) {
    // kfilter.hpp:135-184 draws the points one by one, inserting each into the series before the next Predict; the
    // device draws the whole conditional path at once (carma_simulate: same joint law, O(ny + nsim) work)
    double s2; vecC om; vecD ma;
    params(s2, om, ma);
    std::vector<double> o = flatten(om);
    vecD ysim(time.size());
    if (time.empty()) return ysim;
    check(carma_simulate(series_->handle(), s2, o.data(), ma.data(), (int)om.size(), 1.0, 0.0, time.data(), time.size(),
                         next_run_seed(), 1, ysim.data()), "carma_simulate");
    return ysim;
}

template class KalmanFilter<double>;
template class KalmanFilter<vecC>;

void KalmanFilter1::params(double& sigsqr, vecC& omega, vecD& ma) const {
    sigsqr = sigsqr_;
    omega.assign(1, std::complex<double>(-omega_, 0.0));  // CAR(1): state-space root -omega (kfilter.cpp:19-48)
    ma.assign(1, 1.0);
}

void KalmanFilterp::params(double& sigsqr, vecC& omega, vecD& ma) const {
    sigsqr = sigsqr_;
    omega = omega_;
    if (omega.empty() || omega.size() > CARMA_MAX_P) throw std::invalid_argument("KalmanFilterp: need 1..7 AR roots");
    ma = ma_coefs_;
    ma.resize(omega.size(), 0.0);  // kfilter.hpp:310-312
    if (ma_coefs_.empty()) ma[0] = 1.0;
}

// ---- CARMA_Base ------------------------------------------------------------------------------
CARMA_Base::CARMA_Base(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int kind, int p,
                       int q, double temperature)
    : Parameter<vecD>(track, name, temperature), kind_(kind), p_(p), q_(q),
      series_(std::make_shared<DeviceSeries>(time, y, yerr)) {
    if (p < 1 || p > CARMA_MAX_P) throw std::invalid_argument("order p must be in 1..7");
    value_.assign((size_t)Dimension(), 0.0);  // carpack.hpp:72: value_.set_size(p + 3)
    // carpack.hpp:71: SetPrior(10 sqrt(arma::var(y))) with the N-1 variance
    check(carma_series_default_prior(series_->handle(), 0, &prior_), "carma_series_default_prior");
}

int CARMA_Base::Dimension() const {
    if (kind_ == CARMA_KIND_CAR1) return 4;
    if (kind_ == CARMA_KIND_CARMA) return 3 + p_ + q_;
    if (kind_ == CARMA_KIND_ZCARMA) return 4 + p_;
    return 3 + p_;
}

double CARMA_Base::LogPrior(const vecD& theta) const {
    if ((int)theta.size() != Dimension()) throw std::invalid_argument("theta has the wrong length");
    double out = 0.0;
    check(carma_log_prior(kind_, p_, theta.data(), &prior_, &out), "carma_log_prior");
    return out;
}

double CARMA_Base::LogDensity(vecD theta) {
    if ((int)theta.size() != Dimension()) throw std::invalid_argument("theta has the wrong length");
    double out = 0.0;
    check(carma_loglik_batch(series_->handle(), kind_, p_, q_, &prior_, 1, theta.data(), &out,
                             ignore_prior_ ? CARMA_IGNORE_BOUNDS : 0u), "carma_loglik_batch");
    last_theta_ = theta;
    last_logpost_ = out;
    return out;
}

void CARMA_Base::Save(vecD new_value) {
    // carpack.hpp:90-108 sums the log-posterior over the mean / var the filter holds from the preceding
    // LogDensity(new_value); a different value here means no such call was made, so evaluate it
    const bool cached = new_value == last_theta_;
    const double lp = cached ? last_logpost_ : LogDensity(new_value);
    value_ = new_value;
    log_posterior_ = lp;
}

vecD CARMA_Base::StartingValue() {
    vecD theta((size_t)Dimension());
    double lp = 0.0;
    // a fresh Philox key per call (seeded by set_seed / the clock), the chain index separates the chains of an ensemble
    check(carma_starting_value(series_->handle(), kind_, p_, q_, &prior_, next_run_seed() + start_draws_++, chain_, 1000,
                               theta.data(), &lp), "carma_starting_value");
    last_theta_ = theta;
    last_logpost_ = lp;
    return theta;
}

vecD CARMA_Base::SetStartingValue(vecD init) {
    if ((int)init.size() != Dimension()) {
        std::cout << "WARNING: initial guess wrong length, initializing with prior" << std::endl;
        return StartingValue();
    }
    const double logpost = LogDensity(init);
    if (!std::isfinite(logpost)) {
        std::cout << "WARNING: initial guess yields non-finite likelihood, initializing with prior" << std::endl;
        return StartingValue();
    }
    return init;
}

std::string CARMA_Base::StringValue() {
    std::ostringstream ss;
    for (size_t i = 0; i < value_.size(); i++) ss << (i ? " " : "") << value_[i];
    return ss.str();
}

vecD CARMA_Base::LogDensityBatch(const vecvecD& theta) const {
    const size_t d = (size_t)Dimension();
    std::vector<double> flat(theta.size() * d);
    for (size_t i = 0; i < theta.size(); i++) {
        if (theta[i].size() != d) throw std::invalid_argument("theta row has the wrong length");
        std::copy(theta[i].begin(), theta[i].end(), flat.begin() + i * d);
    }
    vecD out(theta.size());
    check(carma_loglik_batch(series_->handle(), kind_, p_, q_, &prior_, theta.size(), flat.data(), out.data(),
                             ignore_prior_ ? CARMA_IGNORE_BOUNDS : 0u), "carma_loglik_batch");
    return out;
}

bool CARMA_Base::CheckPriorBounds(const vecD& theta) {
    if (ignore_prior_ && kind_ != CARMA_KIND_CAR1) return true;
    double lp = LogDensity(theta);
    return !(std::isinf(lp) && lp < 0);
}

CARMA::CARMA(bool track, std::string name, const vecD& time, const vecD& y, const vecD& yerr, int p, int q,
             double temperature)
    : CARp(track, name, time, y, yerr, CARMA_KIND_CARMA, p, q, temperature) {
    if (!(q < p) || q < 0)  // BOOST_ASSERT_MSG(q < p, ...) carpack.hpp:377
        throw std::invalid_argument("Order of moving average polynomial must be less than order of autoregressive polynomial");
}

vecC CARp::ARRoots(const vecD& theta) const {  // carpack.cpp:137-172
    vecC roots(p_);
    for (int i = 0; i < p_ / 2; i++) {
        double q1 = std::exp(theta[3 + 2 * i]), q2 = std::exp(theta[3 + 2 * i + 1]);
        double disc = q2 * q2 - 4.0 * q1;
        if (disc > 0) {
            roots[2 * i] = std::complex<double>(-0.5 * (q2 + std::sqrt(disc)), 0.0);
            roots[2 * i + 1] = std::complex<double>(-0.5 * (q2 - std::sqrt(disc)), 0.0);
        } else {
            roots[2 * i] = std::complex<double>(-0.5 * q2, -0.5 * std::sqrt(-disc));
            roots[2 * i + 1] = std::conj(roots[2 * i]);
        }
    }
    if (p_ % 2 == 1) roots[p_ - 1] = std::complex<double>(-std::exp(theta[3 + p_ - 1]), 0.0);
    return roots;
}

double CARp::Variance(const vecC& r, const vecD& ma, double sigma, double dt) const {  // carpack.cpp:377-409
    std::complex<double> total(0, 0);
    for (size_t k = 0; k < r.size(); k++) {
        std::complex<double> dp(1, 0), s1(0, 0), s2(0, 0), pw1(1, 0), pw2(1, 0);
        for (size_t l = 0; l < r.size(); l++)
            if (l != k) dp *= (r[l] - r[k]) * (std::conj(r[l]) + r[k]);
        for (size_t l = 0; l < ma.size(); l++) { s1 += ma[l] * pw1; s2 += ma[l] * pw2; pw1 *= r[k]; pw2 *= -r[k]; }
        total += s1 * s2 * std::exp(r[k] * dt) / (-2.0 * r[k].real() * dp);
    }
    return sigma * sigma * total.real();
}

vecD CARp::ExtractMA(const vecD& theta) const {  // carpack.hpp:314, 335: [1, 0, ..., 0] (also what ZCAR evaluates)
    (void)theta;
    vecD ma((size_t)p_, 0.0);
    ma[0] = 1.0;
    return ma;
}

vecD CARMA::ExtractMA(const vecD& theta) const {  // carpack.cpp:522-580 with polycoefs (742-756)
    vecD ma((size_t)p_, 0.0);
    ma[0] = 1.0;
    if (q_ == 0) return ma;
    vecC roots((size_t)q_);
    for (int i = 0; i < q_ / 2; i++) {
        const double q1 = std::exp(theta[3 + p_ + 2 * i]), q2 = std::exp(theta[3 + p_ + 2 * i + 1]);
        const double disc = q2 * q2 - 4.0 * q1;
        if (disc > 0) {
            roots[2 * i] = std::complex<double>(-0.5 * (q2 + std::sqrt(disc)), 0.0);
            roots[2 * i + 1] = std::complex<double>(-0.5 * (q2 - std::sqrt(disc)), 0.0);
        } else {
            roots[2 * i] = std::complex<double>(-0.5 * q2, -0.5 * std::sqrt(-disc));
            roots[2 * i + 1] = std::conj(roots[2 * i]);
        }
    }
    if (q_ % 2 == 1) roots[q_ - 1] = std::complex<double>(-std::exp(theta[3 + p_ + q_ - 1]), 0.0);
    vecC coefs((size_t)q_ + 1, std::complex<double>(0, 0));
    coefs[0] = 1.0;
    for (int i = 0; i < q_; i++)
        for (int j = i + 1; j >= 1; j--) coefs[j] = coefs[j] - roots[i] * coefs[j - 1];
    const double norm = coefs[q_].real();
    for (int i = 0; i <= q_; i++) ma[i] = coefs[q_ - i].real() / norm;
    return ma;
}

vecD ZCARMA::ExtractMA(const vecD& theta) const {  // carpack.cpp:687-698
    const double x = theta[3 + p_];
    const double kn = std::exp(x) / (1.0 + std::exp(x));
    const double kappa = (prior_.kappa_high - prior_.kappa_low) * kn + prior_.kappa_low;
    vecD ma((size_t)p_, 0.0);
    ma[0] = 1.0;
    double binom = 1.0;
    for (int i = 1; i < p_; i++) {
        binom = binom * (double)(p_ - i) / (double)i;
        ma[i] = std::rint(binom) / std::pow(kappa, (double)i);
    }
    return ma;
}

// ---- samplers --------------------------------------------------------------------------------
static void run_pt(CARMA_Base& par, int sample_size, int burnin, int nwalkers, int thin, const vecD& init, uint64_t seed,
                   size_t n_ens, std::vector<vecvecD>& samples, std::vector<vecD>& logposts, std::vector<vecD>& acc,
                   std::vector<vecD>& exch) {
    carma_pt_opts_t o;
    carma_pt_default_opts(&o);
    o.nsamples = sample_size; o.burnin = burnin; o.thin = thin; o.ntemps = nwalkers; o.seed = seed;
    const size_t d = (size_t)par.Dimension();
    std::vector<double> s(n_ens * (size_t)sample_size * d), lp(n_ens * (size_t)sample_size), ar(n_ens * nwalkers), xr(n_ens * nwalkers);
    const double* pinit = nullptr;
    if (init.size() == d) pinit = init.data();
    else if (!init.empty()) std::cout << "WARNING: initial guess wrong length, initializing with prior" << std::endl;  // carpack.cpp:481-484
    std::cout << "Running sampler on the GPU: " << n_ens << " ensemble(s) x " << nwalkers << " temperature(s), "
              << burnin << " burn-in + " << sample_size << " x " << thin << " iterations" << std::endl;
    check(carma_pt_run(par.series()->handle(), par.kind(), par.p(), par.q(), &par.prior(), &o, n_ens, pinit, s.data(),
                       lp.data(), ar.data(), xr.data(), nullptr, nullptr, nullptr), "carma_pt_run");
    samples.assign(n_ens, vecvecD());
    logposts.assign(n_ens, vecD());
    acc.assign(n_ens, vecD());
    exch.assign(n_ens, vecD());
    for (size_t e = 0; e < n_ens; e++) {
        samples[e].resize(sample_size);
        for (int i = 0; i < sample_size; i++)
            samples[e][i].assign(s.begin() + (e * sample_size + i) * d, s.begin() + (e * sample_size + i + 1) * d);
        logposts[e].assign(lp.begin() + e * sample_size, lp.begin() + (e + 1) * sample_size);
        acc[e].assign(ar.begin() + e * nwalkers, ar.begin() + (e + 1) * nwalkers);
        exch[e].assign(xr.begin() + e * nwalkers, xr.begin() + (e + 1) * nwalkers);
    }
    std::cout << "Average RAM Acceptance Rate (coolest chain) is " << acc[0][0] << std::endl;  // steps.cpp:103-106
}

static double population_max_stdev(const vecD& y) {  // carmcmc.cpp:85-89
    double sum = std::accumulate(y.begin(), y.end(), 0.0);
    double mean = sum / y.size();
    double sq = std::inner_product(y.begin(), y.end(), y.begin(), 0.0);
    return 10.0 * std::sqrt(sq / y.size() - mean * mean);
}

std::shared_ptr<CAR1> RunCar1Sampler(int sample_size, int burnin, vecD time, vecD y, vecD yerr, int thin, const vecD& init) {
    std::shared_ptr<CAR1> par = std::make_shared<CAR1>(true, "CAR(1)", time, y, yerr);
    par->SetPrior(population_max_stdev(par->GetTimeSeries()));
    std::vector<vecvecD> s; std::vector<vecD> lp, ar, xr;
    run_pt(*par, sample_size, burnin, 1, thin, init, next_run_seed(), 1, s, lp, ar, xr);
    par->SetSamples(s[0], lp[0]);
    par->accept_rates = ar[0]; par->exchange_rates = xr[0];
    return par;
}

static std::shared_ptr<CARp> make_carma_par(const vecD& time, const vecD& y, const vecD& yerr, int p, int q, bool do_zcarma) {
    std::shared_ptr<CARp> par;
    if (do_zcarma) par = std::make_shared<ZCAR>(true, "ZCAR(p) Parameters", time, y, yerr, p);          // carmcmc.cpp:112
    else if (q == 0) par = std::make_shared<CARp>(true, "CAR(p) Parameters", time, y, yerr, p);          // carmcmc.cpp:105
    else par = std::make_shared<CARMA>(true, "CARMA(p,q) Parameters", time, y, yerr, p, q);              // carmcmc.cpp:108
    par->SetPrior(population_max_stdev(par->GetTimeSeries()));
    return par;
}

std::shared_ptr<CARp> RunCarmaSampler(int sample_size, int burnin, vecD time, vecD y, vecD yerr, int p, int q,
                                      int nwalkers, bool do_zcarma, int thin, const vecD& init) {
    if (!(p > 1)) throw std::invalid_argument("RunCarmaSampler requires p > 1 (use RunCar1Sampler for p == 1)");  // carmcmc.cpp:84
    std::shared_ptr<CARp> par = make_carma_par(time, y, yerr, p, q, do_zcarma);
    std::vector<vecvecD> s; std::vector<vecD> lp, ar, xr;
    run_pt(*par, sample_size, burnin, nwalkers, thin, init, next_run_seed(), 1, s, lp, ar, xr);
    par->SetSamples(s[0], lp[0]);
    par->accept_rates = ar[0]; par->exchange_rates = xr[0];
    return par;
}

std::vector<std::shared_ptr<CARp> > RunCarmaSamplerEnsembles(int n_ensembles, int sample_size, int burnin, vecD time,
                                                             vecD y, vecD yerr, int p, int q, int nwalkers,
                                                             bool do_zcarma, int thin, const vecD& init) {
    if (!(p > 1)) throw std::invalid_argument("RunCarmaSamplerEnsembles requires p > 1");
    std::shared_ptr<CARp> first = make_carma_par(time, y, yerr, p, q, do_zcarma);
    std::vector<vecvecD> s; std::vector<vecD> lp, ar, xr;
    run_pt(*first, sample_size, burnin, nwalkers, thin, init, next_run_seed(), (size_t)n_ensembles, s, lp, ar, xr);
    std::vector<std::shared_ptr<CARp> > out;
    for (int e = 0; e < n_ensembles; e++) {
        // every returned object is a fully functional evaluator bound to the same device series
        std::shared_ptr<CARp> par = (e == 0) ? first : std::make_shared<CARp>(*first);
        par->SetSamples(s[e], lp[e]);
        par->accept_rates = ar[e]; par->exchange_rates = xr[e];
        out.push_back(par);
    }
    return out;
}

}  // namespace carma_host
