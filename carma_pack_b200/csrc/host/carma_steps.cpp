// carma_steps.cpp -- bodies of the generic MCMC classes of carma_steps.hpp (host code; see the header).
#include "carma_steps.hpp"

#include <chrono>
#include <stdexcept>

namespace carma_host {

std::mt19937_64 rng((uint64_t)std::chrono::system_clock::now().time_since_epoch().count());  // random.cpp:20
RandomGenerator RandGen;

double RandomGenerator::normal(double mu, double sigma) { return std::normal_distribution<double>(mu, sigma)(rng); }
double RandomGenerator::uniform(double lo, double hi) { return std::uniform_real_distribution<double>(lo, hi)(rng); }
double RandomGenerator::chisqr(int dof) { return std::chi_squared_distribution<double>((double)dof)(rng); }
double RandomGenerator::tdist(int dof, double mean, double scale) {
    return mean + scale * std::student_t_distribution<double>((double)dof)(rng);
}
double RandomGenerator::scaled_inverse_chisqr(int dof, double ssqr) { return ssqr / chisqr(dof) * (double)dof; }

matD chol_upper(const matD& a) {
    const size_t n = a.size();
    matD r(n, vecD(n, 0.0));
    for (size_t j = 0; j < n; j++) {
        double s = a[j][j];
        for (size_t k = 0; k < j; k++) s -= r[k][j] * r[k][j];
        if (!(s > 0.0)) throw std::runtime_error("chol(): matrix is not positive definite");
        r[j][j] = std::sqrt(s);
        for (size_t i = j + 1; i < n; i++) {
            double t = a[j][i];
            for (size_t k = 0; k < j; k++) t -= r[k][j] * r[k][i];
            r[j][i] = t / r[j][j];
        }
    }
    return r;
}

void CholUpdateR1(matD& L, vecD& v, bool downdate) {  // steps.cpp:111-131
    const double sign = downdate ? -1.0 : 1.0;
    const size_t n = L.size();
    for (size_t k = 0; k < n; k++) {
        const double r = std::sqrt(L[k][k] * L[k][k] + sign * v[k] * v[k]);
        const double c = r / L[k][k];
        const double s = v[k] / L[k][k];
        L[k][k] = r;
        for (size_t j = k + 1; j < n; j++) {
            L[k][j] = (L[k][j] + sign * s * v[j]) / c;
            v[j] = c * v[j] - s * L[k][j];
        }
    }
}

AdaptiveMetro::AdaptiveMetro(Parameter<vecD>& parameter, Proposal<double>& proposal, matD proposal_covar, double target_rate,
                             int maxiter)
    : parameter_(parameter), proposal_(proposal), target_rate_(target_rate), maxiter_(maxiter) {
    gamma_ = 2.0 / 3.0;
    niter_ = 0;
    naccept_ = 0;
    chol_factor_ = chol_upper(proposal_covar);
}

matD AdaptiveMetro::GetCovariance() const {
    const size_t n = chol_factor_.size();
    matD c(n, vecD(n, 0.0));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++)
            for (size_t k = 0; k < n; k++) c[i][j] += chol_factor_[k][i] * chol_factor_[k][j];
    return c;
}

// steps.cpp:36-56.  The proposal's log-density is kept for Save(): LogDensity is the expensive call (one filter run).
bool AdaptiveMetro::Accept(vecD new_value, vecD old_value) {
    (void)old_value;
    last_logdensity_ = parameter_.LogDensity(new_value);
    alpha_ = (last_logdensity_ - parameter_.GetLogDensity()) / parameter_.GetTemperature();
    if (!std::isfinite(alpha_)) {
        alpha_ = 0.0;
        return false;
    }
    const double unif = RandGen.uniform();
    alpha_ = std::min(std::exp(alpha_), 1.0);
    if (unif < alpha_) {
        naccept_++;
        return true;
    }
    return false;
}

void AdaptiveMetro::DoStep() {  // steps.cpp:60-107
    vecD old_value = parameter_.Value();
    const size_t n = old_value.size();
    vecD unit_proposal(n), scaled_proposal(n, 0.0), new_value(n);
    for (size_t i = 0; i < n; i++) unit_proposal[i] = proposal_.Draw(0.0);
    for (size_t j = 0; j < n; j++) {  // chol_factor_.t() * unit_proposal
        double s = 0.0;
        for (size_t k = 0; k <= j; k++) s += chol_factor_[k][j] * unit_proposal[k];
        scaled_proposal[j] = s;
        new_value[j] = old_value[j] + s;
    }
    if (Accept(new_value, old_value)) parameter_.Save(new_value);
    if ((niter_ < maxiter_) && std::isfinite(alpha_)) {
        const double step_size = std::min(1.0, (double)n / std::pow((double)niter_, gamma_));
        double unit_norm = 0.0;
        for (size_t i = 0; i < n; i++) unit_norm += unit_proposal[i] * unit_proposal[i];
        unit_norm = std::sqrt(unit_norm);
        const double f = std::sqrt(step_size * std::fabs(alpha_ - target_rate_)) / unit_norm;
        for (size_t i = 0; i < n; i++) scaled_proposal[i] *= f;
        CholUpdateR1(chol_factor_, scaled_proposal, alpha_ < target_rate_);
    }
    niter_++;
    if (niter_ == maxiter_) std::cout << "Average RAM Acceptance Rate is " << (double)naccept_ / (double)niter_ << std::endl;
}

void Sampler::AddStep(Step* step) {  // samplers.cpp:24-34
    steps_.emplace_back(step);
    if (steps_.back()->ParameterTrack()) {
        std::string par_label = steps_.back()->ParameterLabel();
        tracked_names_.insert(par_label);
        p_tracked_parameters_[par_label] = steps_.back()->GetParPointer();
    }
}

void Sampler::Iterate(int number_of_iterations, bool progress) {  // samplers.cpp:37-54
    (void)progress;
    for (int iter = 0; iter < number_of_iterations; ++iter)
        for (size_t i = 0; i < steps_.size(); ++i) steps_[i]->DoStep();
}

void Sampler::Run(vecD init) {  // samplers.cpp:57-115
    current_iter_ = 0;
    for (const std::string& label : tracked_names_) p_tracked_parameters_[label]->SetSampleSize(sample_size_);
    if (verbose) {
        std::cout << "Running sampler..." << std::endl;
        std::cout << "Number of steps added: " << NumberOfSteps() << std::endl;
        std::cout << "Number of tracked steps added: " << NumberOfTrackedSteps() << std::endl;
        std::cout << "Setting starting values..." << std::endl;
    }
    if (steps_.empty()) return;
    // one initialisation per distinct parameter (the reference initialises per STEP, so a chain that owns a RAM step
    // and an exchange step is drawn twice and the second draw wins: distributionally identical)
    std::set<BaseParameter*> done;
    const size_t npar = static_cast<Parameter<vecD>*>(steps_[0]->GetParPointer())->Value().size();
    const bool useInit = (init.size() == npar) && npar > 0;  // samplers.cpp:76
    if (verbose) std::cout << (useInit ? " Using user-provided values" : " Drawn from priors") << std::endl;
    for (size_t i = 0; i < steps_.size(); ++i) {
        Parameter<vecD>* par = static_cast<Parameter<vecD>*>(steps_[i]->GetParPointer());
        if (!done.insert(par).second) continue;
        if (useInit) par->Save(par->SetStartingValue(init));
        else par->Save(par->StartingValue());
    }
    if (verbose) std::cout << "Burning in... (" << burnin_ << " iterations)" << std::endl;
    Iterate(burnin_, true);
    if (verbose) std::cout << std::endl << "Sampling..." << std::endl;
    for (int i = 0; i < sample_size_; ++i) {
        Iterate(thin_);
        SaveValues();
        ++current_iter_;
    }
}

void Sampler::SaveValues() {  // samplers.cpp:118-124
    for (const std::string& label : tracked_names_) p_tracked_parameters_[label]->AddToSample(current_iter_);
}

}  // namespace carma_host
