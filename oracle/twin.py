"""TEST INFRASTRUCTURE (like everything under oracle/): sensitivity twin for parity checks on states where the reference's own
algorithm is ill-conditioned.

MuJoCo sends cylinder-box / cylinder-cylinder / capsule-cylinder pairs through libccd's MPR (oracle/mjstep_ref.c::mpr_convex,
csrc/b2_mpr.cuh).  For a flat cap resting on a flat face every point under the cap is equally deep, and which one MPR
returns is decided by the last bits of the geom poses: perturbing a pose by 1e-8 moves the contact point across the whole cap
(tests/test_oracle_mpr.py::test_flat_contact_point_can_be_decided_by_rounding).  The arm scene as authored is full of such
contacts (four screws standing on their 2 mm shaft caps, the wrist cylinders lying on the fixture and the table), so a second
implementation -- the fp32 kernel here, or MuJoCo itself on another compiler -- cannot be expected to agree with the oracle
there more closely than the oracle agrees with itself under an fp32-sized perturbation of the state.

`perturbed(...)` returns a copy of an oracle state moved by such a perturbation; a parity test that uses it states its bound
as  base tolerance + SLACK x |oracle(perturbed) - oracle|  per compared entry, i.e. tight wherever the oracle is well
conditioned and as loose as the oracle's own response where it is not.
"""
import numpy as np

SLACK = 4.0
REL = 6e-8      # fp32 unit round-off
ABS = 1e-8      # moves exact zeros (identity quaternions, velocities at rest) off their symmetric configuration


def perturbed(x, rng):
    x = np.asarray(x, np.float64)
    return x * (1.0 + REL * rng.standard_normal(x.shape)) + ABS * rng.standard_normal(x.shape)


def spread(samples, base):
    """max over the ensemble of |sample - base|, element-wise"""
    return np.max(np.abs(np.asarray(samples) - np.asarray(base)[None]), axis=0)
