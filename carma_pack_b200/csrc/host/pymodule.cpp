// pymodule.cpp -- the `_carmcmc` extension module: same Python-visible names as the reference's
// Boost.Python module (src/boost_python_wrapper.cpp:28-102), built with pybind11 on top of the
// reference-named host classes (carma_host.hpp), which call the GPU through the C ABI.
//
//   containers  vecD, vecvecD, vecC, pairD{first,second}                      (wrapper :32-43)
//   classes     CAR1, CARp, CARMA  (+ ZCARMA, which the reference defines but never exports)
//               getLogPrior, getLogDensity, getSamples, GetLogLikes, SetMLE    (wrapper :49-73)
//   functions   run_mcmc_car1 (5-7 args), run_mcmc_carma (8-11 args)           (wrapper :25-26, 76-77)
//   filters     KalmanFilter1, KalmanFilterp: Simulate, Filter, Predict, GetMean, GetVar (wrapper :83-101)
// Extensions (not in the reference): getLogDensityBatch, PredictMany, run_mcmc_carma_ensembles,
// set_seed, and implicit conversion of Python lists / numpy arrays to vecD / vecC.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl_bind.h>

#include "carma_host.hpp"

namespace py = pybind11;
using namespace carma_host;

PYBIND11_MAKE_OPAQUE(vecD);
PYBIND11_MAKE_OPAQUE(vecvecD);
PYBIND11_MAKE_OPAQUE(vecC);

struct PairD {
    double first = 0.0, second = 0.0;
};
template <class K>
static PairD predict_pair(K& k, double t) {
    std::pair<double, double> p = k.Predict(t);
    PairD out;
    out.first = p.first;
    out.second = p.second;
    return out;
}

PYBIND11_MODULE(_carmcmc, m) {
    m.doc() = "B200-native drop-in for carma_pack's _carmcmc extension (GPU Kalman log-likelihood + PT-MCMC)";

    py::bind_vector<vecD>(m, "vecD", py::buffer_protocol());
    py::bind_vector<vecvecD>(m, "vecvecD");
    py::bind_vector<vecC>(m, "vecC");
    py::implicitly_convertible<py::list, vecD>();
    py::implicitly_convertible<py::tuple, vecD>();
    py::implicitly_convertible<py::array, vecD>();
    py::implicitly_convertible<py::list, vecC>();
    py::implicitly_convertible<py::array, vecC>();

    // std::pair has a built-in tuple caster in pybind11, so the reference's pairD class (wrapper :41-43)
    // is a small struct of its own with the same two read/write members.
    py::class_<PairD>(m, "pairD")
        .def(py::init<>())
        .def_readwrite("first", &PairD::first)
        .def_readwrite("second", &PairD::second)
        .def("__iter__", [](const PairD& p) { return py::iter(py::make_tuple(p.first, p.second)); })
        .def("__repr__", [](const PairD& p) { return "pairD(" + std::to_string(p.first) + ", " + std::to_string(p.second) + ")"; });

    // ---- carpack.hpp ------------------------------------------------------------------------
    py::class_<CARMA_Base, std::shared_ptr<CARMA_Base> >(m, "CARMA_Base")
        .def("getLogPrior", &CARMA_Base::getLogPrior)
        .def("getLogDensity", &CARMA_Base::getLogDensity)
        .def("getLogDensityBatch", &CARMA_Base::LogDensityBatch)
        .def("getSamples", &CARMA_Base::getSamples)
        .def("GetLogLikes", &CARMA_Base::GetLogLikes)
        .def("SetMLE", &CARMA_Base::SetMLE)
        .def("SetPrior", &CARMA_Base::SetPrior)
        .def("CheckPriorBounds", &CARMA_Base::CheckPriorBounds)
        .def("Dimension", &CARMA_Base::Dimension)
        .def_readonly("accept_rates", &CARMA_Base::accept_rates)
        .def_readonly("exchange_rates", &CARMA_Base::exchange_rates);

    py::class_<CAR1, CARMA_Base, std::shared_ptr<CAR1> >(m, "CAR1")
        .def(py::init<bool, std::string, vecD, vecD, vecD, double>(), py::arg("track"), py::arg("name"), py::arg("time"),
             py::arg("y"), py::arg("yerr"), py::arg("temperature") = 1.0);

    py::class_<CARp, CARMA_Base, std::shared_ptr<CARp> >(m, "CARp")
        .def(py::init<bool, std::string, vecD, vecD, vecD, int, double>(), py::arg("track"), py::arg("name"),
             py::arg("time"), py::arg("y"), py::arg("yerr"), py::arg("p"), py::arg("temperature") = 1.0)
        .def("ARRoots", &CARp::ARRoots)
        .def("Variance", &CARp::Variance, py::arg("alpha_roots"), py::arg("ma_coefs"), py::arg("sigma"), py::arg("dt") = 0.0);

    py::class_<CARMA, CARp, std::shared_ptr<CARMA> >(m, "CARMA")
        .def(py::init<bool, std::string, vecD, vecD, vecD, int, int, double>(), py::arg("track"), py::arg("name"),
             py::arg("time"), py::arg("y"), py::arg("yerr"), py::arg("p"), py::arg("q"), py::arg("temperature") = 1.0);

    py::class_<ZCAR, CARp, std::shared_ptr<ZCAR> >(m, "ZCAR")
        .def(py::init<bool, std::string, vecD, vecD, vecD, int, double>(), py::arg("track"), py::arg("name"),
             py::arg("time"), py::arg("y"), py::arg("yerr"), py::arg("p"), py::arg("temperature") = 1.0);

    py::class_<ZCARMA, CARp, std::shared_ptr<ZCARMA> >(m, "ZCARMA")
        .def(py::init<bool, std::string, vecD, vecD, vecD, int, double>(), py::arg("track"), py::arg("name"),
             py::arg("time"), py::arg("y"), py::arg("yerr"), py::arg("p"), py::arg("temperature") = 1.0)
        .def("SetKappaBounds", &ZCARMA::SetKappaBounds);

    // ---- carmcmc.hpp ------------------------------------------------------------------------
    m.def("run_mcmc_car1", &RunCar1Sampler, py::arg("sample_size"), py::arg("burnin"), py::arg("time"), py::arg("y"),
          py::arg("yerr"), py::arg("thin") = 1, py::arg("init") = vecD(),
          py::call_guard<py::gil_scoped_release>());
    m.def("run_mcmc_carma", &RunCarmaSampler, py::arg("sample_size"), py::arg("burnin"), py::arg("time"), py::arg("y"),
          py::arg("yerr"), py::arg("p"), py::arg("q"), py::arg("nwalkers"), py::arg("do_zcarma") = false,
          py::arg("thin") = 1, py::arg("init") = vecD(), py::call_guard<py::gil_scoped_release>());
    m.def("run_mcmc_carma_ensembles", &RunCarmaSamplerEnsembles, py::arg("n_ensembles"), py::arg("sample_size"),
          py::arg("burnin"), py::arg("time"), py::arg("y"), py::arg("yerr"), py::arg("p"), py::arg("q"),
          py::arg("nwalkers"), py::arg("do_zcarma") = false, py::arg("thin") = 1, py::arg("init") = vecD(),
          py::call_guard<py::gil_scoped_release>());
    m.def("set_seed", &set_seed, "Seed of the Philox streams used by run_mcmc_* (reference: rng(time(NULL)))");

    // ---- kfilter.hpp ------------------------------------------------------------------------
    py::class_<KalmanFilter1, std::shared_ptr<KalmanFilter1> >(m, "KalmanFilter1")
        .def(py::init<vecD, vecD, vecD>())
        .def(py::init<vecD, vecD, vecD, double, double>())
        .def("Simulate", &KalmanFilter1::Simulate)
        .def("Filter", &KalmanFilter1::Filter)
        .def("Predict", &predict_pair<KalmanFilter1>)
        .def("PredictMany", [](KalmanFilter1& k, const vecD& t) { vecD a, b; k.PredictMany(t, a, b); return py::make_tuple(a, b); })
        .def("GetMean", &KalmanFilter1::GetMeanSvec)
        .def("GetVar", &KalmanFilter1::GetVarSvec);

    py::class_<KalmanFilterp, std::shared_ptr<KalmanFilterp> >(m, "KalmanFilterp")
        .def(py::init<vecD, vecD, vecD>())
        .def(py::init<vecD, vecD, vecD, double, vecC, vecD>())
        .def("Simulate", &KalmanFilterp::Simulate)
        .def("Filter", &KalmanFilterp::Filter)
        .def("Predict", &predict_pair<KalmanFilterp>)
        .def("PredictMany", [](KalmanFilterp& k, const vecD& t) { vecD a, b; k.PredictMany(t, a, b); return py::make_tuple(a, b); })
        .def("GetMean", &KalmanFilterp::GetMeanSvec)
        .def("GetVar", &KalmanFilterp::GetVarSvec);
}
