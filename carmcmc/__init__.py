"""Drop-in alias of the reference's package name: `import carmcmc as cm` keeps working.

Mirrors src/carmcmc/__init__.py:1-4 of brandonckelly/carma_pack (`from ._carmcmc import *`, then
CarmaModel, CarmaSample, Car1Sample, power_spectrum, carma_variance, carma_process, get_ar_roots,
MCMCSample), with every name served by the B200 path in carma_pack_b200.
"""
from carma_pack_b200._carmcmc import *  # noqa: F401,F403  vecD, vecvecD, vecC, pairD, CAR1, CARp, CARMA, run_mcmc_*, KalmanFilter*
from carma_pack_b200 import _carmcmc  # noqa: F401
from carma_pack_b200.carma_pack import (CarmaModel, CarmaSample, Car1Sample, MCMCSample, power_spectrum,  # noqa: F401
                                        carma_variance, carma_process, car1_process, get_ar_roots)
