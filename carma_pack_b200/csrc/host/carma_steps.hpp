// carma_steps.hpp -- the reference's generic MCMC machinery with its own class and method names, so that C++ code
// written against carma_pack's step / sampler API links against this library unchanged:
//
//   BaseParameter, Parameter<T>, Ensemble<T>      include/parameters.hpp:32-208
//   Proposal<T>, NormalProposal, StudentProposal  include/proposals.hpp:34-75, proposals.cpp:20-28
//   RandomGenerator (normal, tdist, uniform, ...) include/random.hpp:27-77, random.cpp:76-186
//   Step, AdaptiveMetro, CholUpdateR1             include/steps.hpp:44-71, 218-288; steps.cpp:24-131
//   ExchangeStep<V, P>                            include/steps.hpp:291-396
//   MCMCOptions, Sampler                          include/samplers.hpp:39-145; samplers.cpp:24-124
//
// These classes drive ONE chain step by step from the host; the only work that goes to the GPU is the
// Parameter<>::LogDensity call of the parameter they are given (CARMA_Base::LogDensity -> carma_loglik_batch).  They
// exist for drop-in compatibility and for the reference's own step-level tests (cpp_tests/carma_unit_tests.cpp:783-911).
// The production path is RunCarmaSampler / carma_pt_run: the same algorithm, every chain of every ensemble, in one
// kernel launch with no host round trip per iteration.
// arma::vec / arma::mat are std::vector<double> / row-major std::vector<std::vector<double>> here (no Armadillo).
#pragma once
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace carma_host {

typedef std::vector<double> vecD;
typedef std::vector<std::vector<double> > matD;  // square, row-major

// ---- random.hpp: one global generator, as in the reference (random.cpp:20) -----------------------------------
extern std::mt19937_64 rng;
class RandomGenerator {
public:
    void SetSeed(unsigned long seed) { rng.seed(seed); }
    double normal(double mu = 0.0, double sigma = 1.0);
    double uniform(double lo = 0.0, double hi = 1.0);
    double chisqr(int dof);
    double tdist(int dof, double mean = 0.0, double scale = 1.0);          // random.cpp:158-164
    double scaled_inverse_chisqr(int dof, double ssqr);                    // random.cpp:180-186
};
extern RandomGenerator RandGen;

// ---- parameters.hpp -----------------------------------------------------------------------------------------
class BaseParameter {
public:
    BaseParameter() {}
    BaseParameter(bool track, std::string label, double temperature = 1.0)
        : track_(track), label_(label), temperature_(temperature) {}
    virtual ~BaseParameter() {}
    double GetLogDensity() { return log_posterior_; }
    void SetLogDensity(double logpost) { log_posterior_ = logpost; }
    double GetTemperature() const { return temperature_; }
    virtual std::string StringValue() { return " "; }
    bool Track() const { return track_; }
    void SetTracking(bool track) { track_ = track; }
    std::string Label() const { return label_; }
    virtual void SetSampleSize(int sample_size) = 0;
    virtual void AddToSample(int current_iter) = 0;

protected:
    bool track_ = true;
    std::string label_;
    double temperature_ = 1.0;
    double log_posterior_ = 0.0;
};

template <class ParValueType>
class Parameter : public BaseParameter {
public:
    Parameter() {}
    Parameter(bool track, std::string label, double temperature = 1.0) : BaseParameter(track, label, temperature) {}
    virtual ParValueType StartingValue() = 0;
    virtual ParValueType SetStartingValue(ParValueType init) = 0;
    virtual double LogDensity(ParValueType value) { (void)value; return 0.0; }
    virtual ParValueType RandomPosterior() { return ParValueType(); }
    ParValueType Value() { return value_; }
    virtual void Save(ParValueType new_value) {
        value_ = new_value;
        log_posterior_ = LogDensity(new_value);
    }
    void SetSampleSize(int sample_size) override {
        samples_.resize(sample_size);
        logposts_.resize(sample_size);
    }
    void AddToSample(int current_iter) override {
        samples_[current_iter] = value_;
        logposts_[current_iter] = log_posterior_;
    }
    void AddToSample(int current_iter, ParValueType value, double logpost) {
        samples_[current_iter] = value;
        logposts_[current_iter] = logpost;
    }
    std::vector<ParValueType> GetSamples() const { return samples_; }
    std::vector<double> GetLogLikes() const { return logposts_; }

protected:
    ParValueType value_;
    std::vector<ParValueType> samples_;
    std::vector<double> logposts_;
};

// boost::ptr_vector semantics: the ensemble owns the objects added to it
template <class EnsembleType>
class Ensemble {
public:
    Ensemble() {}
    void AddObject(EnsembleType* pObject) { the_objects_.emplace_back(pObject); }
    int size() const { return (int)the_objects_.size(); }
    EnsembleType& operator[](const int index) { return *the_objects_[index]; }
    EnsembleType const& operator[](const int index) const { return *the_objects_[index]; }
    std::vector<std::unique_ptr<EnsembleType> > the_objects_;
};

// ---- proposals.hpp ------------------------------------------------------------------------------------------
template <typename ProposalType>
class Proposal {
public:
    virtual ~Proposal() {}
    virtual ProposalType Draw(ProposalType starting_value) = 0;
    virtual double LogDensity(ProposalType new_value, ProposalType starting_value) = 0;
};
class NormalProposal : public Proposal<double> {
public:
    NormalProposal() {}
    explicit NormalProposal(double standard_deviation) : standard_deviation_(standard_deviation) {}
    double Draw(double starting_value) override { return RandGen.normal(starting_value, standard_deviation_); }
    double LogDensity(double, double) override { return 0.0; }  // symmetric
private:
    double standard_deviation_ = 1.0;
};
class StudentProposal : public Proposal<double> {
public:
    StudentProposal() {}
    StudentProposal(double dof, double scale) : dof_(dof), scale_(scale) {}
    double Draw(double starting_value) override { return RandGen.tdist((int)dof_, starting_value, scale_); }  // proposals.cpp:26-28
    double LogDensity(double, double) override { return 0.0; }  // symmetric
private:
    double dof_ = 8.0, scale_ = 1.0;
};

// ---- steps.hpp ----------------------------------------------------------------------------------------------
class Step {
public:
    virtual ~Step() {}
    virtual void DoStep() = 0;
    virtual std::string ParameterLabel() { return " "; }
    virtual std::string ParameterValue() { return " "; }
    virtual bool ParameterTrack() { return true; }
    virtual BaseParameter* GetParPointer() = 0;
};

// upper-triangular Cholesky factor R of a symmetric positive-definite matrix, A = R^T R (arma::chol)
matD chol_upper(const matD& a);
// rank-1 update / downdate of the upper factor (steps.cpp:111-131); v is overwritten
void CholUpdateR1(matD& L, vecD& v, bool downdate);

// Robust Adaptive Metropolis (Vihola 2012): steps.hpp:218-288, steps.cpp:24-107
class AdaptiveMetro : public Step {
public:
    AdaptiveMetro(Parameter<vecD>& parameter, Proposal<double>& proposal, matD proposal_covar, double target_rate, int maxiter);
    std::string ParameterLabel() override { return parameter_.Label(); }
    std::string ParameterValue() override { return parameter_.StringValue(); }
    void SetTargetRate(double target_rate) { target_rate_ = target_rate; }
    void SetDecayRate(double gamma) { gamma_ = gamma; }
    bool Accept(vecD new_value, vecD old_value);
    void DoStep() override;
    double GetMetroRatio() const { return alpha_; }
    double GetAcceptRate() const { return (double)naccept_ / (double)niter_; }
    matD GetCovariance() const;  // chol_factor_.t() * chol_factor_
    bool ParameterTrack() override { return parameter_.Track(); }
    BaseParameter* GetParPointer() override { return &parameter_; }

private:
    Parameter<vecD>& parameter_;
    Proposal<double>& proposal_;
    matD chol_factor_;
    double gamma_, target_rate_;
    int niter_, naccept_, maxiter_;
    double alpha_ = 0.0;
    double last_logdensity_ = 0.0;
};

// Parallel-tempering exchange between ensemble[parameter_index] and ensemble[parameter_index - 1]: steps.hpp:291-396
template <class ParValueType, class ParameterType>
class ExchangeStep : public Step {
public:
    ExchangeStep(Parameter<ParValueType>& parameter, int parameter_index, Ensemble<ParameterType>& ensemble, int report_iter = -1)
        : parameter_(parameter), parameter_index_(parameter_index), ensemble_(ensemble), report_iter_(report_iter) {
        if (!(parameter_index_ > 0)) throw std::invalid_argument("ExchangeStep: parameter_index must be > 0");
    }
    std::string ParameterLabel() override { return parameter_.Label(); }
    std::string ParameterValue() override { return parameter_.StringValue(); }
    void DoStep() override {
        const double this_logpost = parameter_.GetLogDensity();
        const double this_temperature = parameter_.GetTemperature();
        ParameterType& other = ensemble_[parameter_index_ - 1];
        const double other_logpost = other.GetLogDensity();
        const double other_temperature = other.GetTemperature();
        double alpha = 1.0 / this_temperature * (other_logpost - this_logpost) + 1.0 / other_temperature * (this_logpost - other_logpost);
        const double unif = RandGen.uniform();
        alpha = std::min(std::exp(alpha), 1.0);  // a NaN stays a NaN here (std::min returns its first argument) ...
        if (!std::isfinite(alpha)) alpha = 0.0;  // ... and is rejected (steps.hpp:334-337)
        if (unif < alpha) {
            // swap the values and the cached log-posteriors; no log-density is recomputed (steps.hpp:342-350)
            ParValueType this_theta = parameter_.Value();
            SwapIn(parameter_, other.Value(), other_logpost);
            SwapIn(other, this_theta, this_logpost);
            naccept_++;
        }
        niter_++;
        if (niter_ == report_iter_) Report();
        alpha_ = alpha;
    }
    void Report() {
        std::cout << "Average Exchange Acceptance Rate Since Last Report: " << (double)naccept_ / (double)niter_ << std::endl;
        niter_ = 0;
        naccept_ = 0;
    }
    double GetMetroRatio() const { return alpha_; }
    bool ParameterTrack() override { return parameter_.Track(); }
    BaseParameter* GetParPointer() override { return &parameter_; }

private:
    // Parameter<>::Save would recompute the log-density (one GPU evaluation per swapped chain) only to have it
    // overwritten by SetLogDensity, as in the reference; SetValue avoids that when the parameter offers it.
    template <class Q>
    static auto SwapIn(Q& par, const ParValueType& v, double lp) -> decltype(par.SetValue(v, lp), void()) { par.SetValue(v, lp); }
    static void SwapIn(Parameter<ParValueType>& par, const ParValueType& v, double lp, ...) { par.Save(v); par.SetLogDensity(lp); }

    Parameter<ParValueType>& parameter_;
    int parameter_index_;
    Ensemble<ParameterType>& ensemble_;
    int report_iter_;
    int niter_ = 0, naccept_ = 0;
    double alpha_ = 0.0;
};

// ---- samplers.hpp -------------------------------------------------------------------------------------------
struct MCMCOptions {
    int sample_size = 0, thin = 1, burnin = 0, chains = 1;
    std::string data_file, out_file;
    int getSampleSize() { return sample_size; }
    void setSampleSize(int s) { sample_size = s; }
    int getThin() { return thin; }
    void setThin(int t) { thin = t; }
    int getBurnin() { return burnin; }
    void setBurnin(int b) { burnin = b; }
    int getChains() { return chains; }
    void setChains(int c) { chains = c; }
    std::string getDataFileName() { return data_file; }
    void setDataFileName(std::string s) { data_file = s; }
    std::string getOutFileName() { return out_file; }
    void setOutFileName(std::string s) { out_file = s; }
};

// samplers.hpp:77-145, samplers.cpp:24-124.  Like the reference's ptr_vector, the sampler owns the steps added to it.
class Sampler {
public:
    Sampler(int sample_size, int burnin, int thin = 1) : sample_size_(sample_size), burnin_(burnin), thin_(thin) {}
    explicit Sampler(MCMCOptions& options) : sample_size_(options.sample_size), burnin_(options.burnin), thin_(options.thin) {}
    void AddStep(Step* step);
    void Iterate(int number_of_iterations, bool progress = false);
    void Run(vecD init = vecD());
    void SaveValues();
    int NumberOfSteps() const { return (int)steps_.size(); }
    int NumberOfTrackedSteps() const { return (int)tracked_names_.size(); }
    std::set<std::string> GetTrackedNames() const { return tracked_names_; }
    std::map<std::string, BaseParameter*> GetTrackedParams() const { return p_tracked_parameters_; }
    bool verbose = true;  // the reference prints its progress to std::cout

private:
    int sample_size_, burnin_, thin_, current_iter_ = 0;
    std::vector<std::unique_ptr<Step> > steps_;
    std::set<std::string> tracked_names_;
    std::map<std::string, BaseParameter*> p_tracked_parameters_;
};

}  // namespace carma_host
