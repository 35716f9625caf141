"""WHY parity is asserted bit-for-bit (and made possible by include/mc_math.h): the reference algorithm is ill-conditioned.

A ONE-ULP perturbation of the air-temperature forcing, fed to the SAME oracle, moves hourly soil temperatures by more than
the north-star tolerance (1e-3 K) at a visible fraction of (hour, node) entries while the median moves by ~1e-7 K: the adaptive
step accept/reject, the Float32 step-size controller, the 1 % Obukhov stopping rule and the loose stomatal Newton exit are
discontinuities that amplify last-bit differences.  Two implementations whose libm differ in the last bit can therefore only
agree statistically; ours share one bit-reproducible math kernel and agree exactly (test_kernel_hostemu.py, test_gpu_parity.py)."""
import copy

import numpy as np

import configs
from oracle.oracle import OracleEngine


def test_one_ulp_of_forcing_exceeds_the_tolerance_somewhere():
    p, _ = configs.config2_daily(12)
    a, _ = configs.run_engine(OracleEngine, p)
    q = copy.deepcopy(p)
    h = q.inputs.environment_hourly
    h.reference_temperature = np.nextafter(np.asarray(h.reference_temperature, float), np.inf)
    b, _ = configs.run_engine(OracleEngine, q)
    d = np.abs(a.soil_temperature - b.soil_temperature)
    assert np.median(d) < 1e-4          # well-conditioned almost everywhere ...
    assert d.max() > 1e-6               # ... but last-bit differences do not stay in the last bit
    # (observed on this fixture: max ~1e-2 K, 0.2 % of entries above 1e-3 K; see DESIGN.md "Parity")
