// C++ test of the reference-named class API (csrc/host/carma_host.hpp, carma_steps.hpp), written the way the
// reference's own Catch tests use it: cpp_tests/carma_unit_tests.cpp:783-911 (CAR1/logpost_test, CARMA/logpost_test:
// StartingValue / Save / RAM.DoStep / GetLogDensity / LogDensity(Value()) / hand-rolled sum over an independent
// KalmanFilterp) and the sampler assembly of src/carmcmc.cpp:97-162 (Ensemble, AdaptiveMetro, ExchangeStep, Sampler).
// Built and run by tests/test_gpu_class_api.py on the GPU box; prints "ok: ..." lines and exits non-zero on failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>

#include "carma_host.hpp"

using namespace carma_host;

static int failures = 0;
#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

static matD eye(size_t n, double v = 1.0) {
    matD m(n, vecD(n, 0.0));
    for (size_t i = 0; i < n; i++) m[i][i] = v;
    return m;
}

// CARMA/logpost_test (carma_unit_tests.cpp:847-911): after every RAM step the cached log-posterior equals
// LogDensity(Value()) and the value summed by hand over an independent Kalman filter + LogPrior, to 1e-10 (abs in the
// reference; relative to max(1,|lp|) here because the three numbers come out of different kernels)
static void carma_logpost_test() {
    const int ny = 100, p = 4, q = 1, niter = 200;
    std::mt19937_64 gen(123456);
    std::normal_distribution<double> nrm(0.0, 1.0);
    vecD time(ny), y(ny), ysig(ny, 0.01);
    for (int i = 0; i < ny; i++) { time[i] = 100.0 * i / (ny - 1); y[i] = 2.0 + nrm(gen); }
    CARMA car_test(true, "CARMA(4,1)", time, y, ysig, p, q);
    double mean = 0, ss = 0;
    for (double v : y) mean += v / ny;
    for (double v : y) ss += (v - mean) * (v - mean);
    car_test.SetPrior(10.0 * std::sqrt(ss / (ny - 1)));

    StudentProposal tUnit(8.0, 1.0);
    AdaptiveMetro RAM(car_test, tUnit, eye(p + 3 + q, 1e-4), 0.4, niter + 1);
    car_test.Save(car_test.StartingValue());   // what Sampler::Run / the reference's RAM.Start() do
    CHECK(std::isfinite(car_test.GetLogDensity()));

    KalmanFilterp Kfilter(time, y, ysig);
    int neq = 0, naccepted = 0;
    vecD prev = car_test.Value();
    for (int i = 0; i < niter; i++) {
        RAM.DoStep();
        const double stored = car_test.GetLogDensity();
        vecD theta = car_test.Value();
        if (theta != prev) naccepted++;
        prev = theta;
        const double computed = car_test.LogDensity(theta);
        // by hand, as the reference's test does it
        vecD ma = car_test.ExtractMA(theta);
        vecC ar = car_test.ARRoots(theta);
        const double sigsqr = theta[0] * theta[0] / car_test.Variance(ar, ma, 1.0);
        vecD ycent(ny), yerr_scaled(ny);
        for (int j = 0; j < ny; j++) { ycent[j] = y[j] - theta[2]; yerr_scaled[j] = std::sqrt(theta[1]) * ysig[j]; }
        KalmanFilterp kf(time, ycent, yerr_scaled, sigsqr, ar, ma);
        kf.Filter();
        double by_hand = 0.0;
        for (int j = 0; j < ny; j++) by_hand += -0.5 * std::log(kf.var[j]) - 0.5 * (ycent[j] - kf.mean[j]) * (ycent[j] - kf.mean[j]) / kf.var[j];
        by_hand += car_test.LogPrior(theta);
        const double tol = 1e-9 * std::max(1.0, std::fabs(stored));
        if (std::fabs(computed - stored) > 1e-10 * std::max(1.0, std::fabs(stored)) || std::fabs(by_hand - stored) > tol) {
            if (neq < 3) std::printf("  mismatch at step %d: stored %.12f computed %.12f by hand %.12f\n", i, stored, computed, by_hand);
            neq++;
        }
    }
    CHECK(neq == 0);
    CHECK(naccepted > 5);
    CHECK(RAM.GetAcceptRate() > 0.02 && RAM.GetAcceptRate() < 0.98);
    matD cov = RAM.GetCovariance();
    CHECK(cov.size() == (size_t)(p + 3 + q) && cov[0][0] > 0);
    std::printf("ok: CARMA/logpost_test, %d RAM steps, %d accepted, acceptance rate %.3f\n", niter, naccepted, RAM.GetAcceptRate());
}

// CAR1/logpost_test (carma_unit_tests.cpp:783-845)
static void car1_logpost_test() {
    const int ny = 100, niter = 200;
    std::mt19937_64 gen(7);
    std::normal_distribution<double> nrm(0.0, 1.0);
    vecD time(ny), y(ny), ysig(ny, 0.01);
    for (int i = 0; i < ny; i++) { time[i] = 100.0 * i / (ny - 1); y[i] = 2.0 + nrm(gen); }
    CAR1 car1_test(true, "CAR(1)", time, y, ysig);
    StudentProposal tUnit(8.0, 1.0);
    AdaptiveMetro RAM(car1_test, tUnit, eye(4, 1e-2), 0.4, niter + 1);
    car1_test.Save(car1_test.StartingValue());
    int neq = 0;
    for (int i = 0; i < niter; i++) {
        RAM.DoStep();
        const double stored = car1_test.GetLogDensity();
        const double computed = car1_test.LogDensity(car1_test.Value());
        if (std::fabs(stored - computed) > 1e-10 * std::max(1.0, std::fabs(stored))) neq++;
    }
    CHECK(neq == 0);
    std::printf("ok: CAR1/logpost_test\n");
}

// The sampler of RunCarmaSampler assembled by hand from the classes (carmcmc.cpp:92-162): a ladder of tempered chains,
// RAM + exchange steps hottest -> coolest, Sampler::Run; stored log-posteriors equal recomputed ones
// (CARMA/logpost_test_mcmc, carma_unit_tests.cpp:1068-1113: 1e-8 relative).
static void hand_built_sampler_test() {
    const int ny = 120, p = 3, q = 1, nwalkers = 4, sample_size = 60, burnin = 80;
    std::mt19937_64 gen(99);
    std::normal_distribution<double> nrm(0.0, 1.0);
    vecD time(ny), y(ny), ysig(ny, 0.1);
    double t = 0;
    for (int i = 0; i < ny; i++) { t += 0.5 + std::fabs(nrm(gen)); time[i] = t; y[i] = (i ? 0.8 * y[i - 1] : 0.0) + nrm(gen); }
    Ensemble<CARMA> ensemble;
    for (int i = 0; i < nwalkers; i++) {
        const double temp = std::exp(std::log(100.0) * i / (nwalkers - 1.0));   // carmcmc.cpp:92-95
        ensemble.AddObject(new CARMA(i == 0, "CARMA(3,1) Parameters", time, y, ysig, p, q, temp));
        ensemble[i].SetChainIndex((uint32_t)i);
    }
    CHECK(ensemble.size() == nwalkers);
    StudentProposal tUnit(8.0, 1.0);
    Sampler sampler(sample_size, burnin, 1);
    sampler.verbose = false;
    std::vector<AdaptiveMetro*> rams;
    for (int i = nwalkers - 1; i > 0; i--) {                                    // carmcmc.cpp:147-157
        rams.push_back(new AdaptiveMetro(ensemble[i], tUnit, eye(p + q + 3, 1e-4), 0.25, burnin));
        sampler.AddStep(rams.back());
        sampler.AddStep(new ExchangeStep<vecD, CARMA>(ensemble[i], i, ensemble, burnin));
    }
    sampler.AddStep(new AdaptiveMetro(ensemble[0], tUnit, eye(p + q + 3, 1e-4), 0.25, burnin));
    CHECK(sampler.NumberOfSteps() == 2 * nwalkers - 1);
    CHECK(sampler.NumberOfTrackedSteps() == 1);   // one label tracked (chain 0; hotter chains are untracked)
    sampler.Run();
    std::vector<vecD> samples = ensemble[0].GetSamples();
    vecD logposts = ensemble[0].GetLogLikes();
    CHECK((int)samples.size() == sample_size && (int)logposts.size() == sample_size);
    vecD recomputed = ensemble[0].LogDensityBatch(samples);
    int bad = 0, distinct = 0;
    for (int i = 0; i < sample_size; i++) {
        if (std::fabs(recomputed[i] - logposts[i]) > 1e-8 * std::max(1.0, std::fabs(logposts[i]))) bad++;
        if (i > 0 && samples[i] != samples[i - 1]) distinct++;
    }
    CHECK(bad == 0);
    CHECK(distinct > 3);
    // hot chains kept their own temperatures and finite cached log-posteriors
    for (int i = 0; i < nwalkers; i++) CHECK(std::isfinite(ensemble[i].GetLogDensity()));
    CHECK(ensemble[nwalkers - 1].GetTemperature() > 99.0 && ensemble[0].GetTemperature() == 1.0);
    std::printf("ok: hand-built PT sampler (%d chains, %d steps), %d distinct consecutive samples\n", nwalkers,
                sampler.NumberOfSteps(), distinct);
}

// CholUpdateR1 against a recomputed factor (steps.cpp:111-131)
static void chol_update_test() {
    matD a = {{4.0, 1.0, 0.5}, {1.0, 3.0, 0.2}, {0.5, 0.2, 2.0}};
    matD r = chol_upper(a);
    vecD v = {0.3, -0.2, 0.5}, v0 = v;
    CholUpdateR1(r, v, false);
    matD a2 = a;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a2[i][j] += v0[i] * v0[j];
    matD r2 = chol_upper(a2);
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) CHECK(std::fabs(r[i][j] - r2[i][j]) < 1e-12);
    v = v0;
    CholUpdateR1(r, v, true);   // and back
    matD r0 = chol_upper(a);
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) CHECK(std::fabs(r[i][j] - r0[i][j]) < 1e-12);
    std::printf("ok: CholUpdateR1 update / downdate\n");
}

int main() {
    set_seed(20261017);
    RandGen.SetSeed(123456);   // carma_unit_tests.cpp:51-53
    chol_update_test();
    car1_logpost_test();
    carma_logpost_test();
    hand_built_sampler_test();
    if (failures) { std::printf("%d check(s) FAILED\n", failures); return 1; }
    std::printf("all class-API checks passed\n");
    return 0;
}
