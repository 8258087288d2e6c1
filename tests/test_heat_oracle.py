"""oracle/heat_oracle.py (c/ch5/heat.c restated) against the reference's goldens and against itself."""
import json
import os

import numpy as np

from oracle import heat_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "heat_goldens.json")))


def test_goldens_are_the_reference_files():
    for n in (1, 2):
        ref = "/root/reference/c/ch5/output/heat.test%d" % n
        if os.path.exists(ref):
            assert open(ref).read().splitlines() == GOLD["heat.test%d" % n]["lines"]


def test_golden_heat_test2_adaptive_rk3bs_verbatim():
    """c/ch5/makefile:47: -da_refine 1 -ts_monitor -ts_type rk -ts_max_time 0.01.  Every digit of the step sequence."""
    _, lines = ho.heat(refine=1, ts_type="rk", tmax=0.01)
    assert lines == GOLD["heat.test2"]["lines"]
    # the controller's exponent is 1/order of the SCHEME (3): the embedded order (2) must miss the golden
    u0 = np.zeros((8, 9))
    _, l2, _ = ho.rk3bs(ho.rhs, u0, 0.001, 0.01, order=2)
    assert l2[1] == "1 TS dt 0.00359127 time 0.001" and l2[1] != GOLD["heat.test2"]["lines"][2]


def test_golden_heat_test1_backward_euler_verbatim():
    _, lines = ho.heat(refine=1, ts_type="beuler")
    assert lines == GOLD["heat.test1"]["lines"]


def test_jacobian_is_the_derivative_of_the_rhs_and_energy_is_conserved():
    rng = np.random.default_rng(3)
    for mx, my in ((9, 8), (17, 12)):
        u, v = rng.standard_normal((my, mx)), rng.standard_normal((my, mx))
        J = ho.jacobian(mx, my, 0.7)
        np.testing.assert_allclose((ho.rhs(u + v, 0.7) - ho.rhs(u, 0.7)).ravel(), J @ v.ravel(), rtol=1e-11, atol=1e-9)
        assert abs(J - ho.jacobian(mx, my, 0.7)).max() == 0.0
    # heat.c's help text: "Energy is conserved (for these particular conditions/source)": d/dt of the discrete integral is
    # the integral of f plus the boundary flux of gamma, both zero over a period in y
    u, lines = ho.heat(refine=2, ts_type="beuler", tmax=0.02, monitor_energy=True)
    e = [float(l.split()[2]) for l in lines if "energy" in l]
    assert len(e) == 21 and max(abs(x) for x in e) < 1e-15 and np.max(np.abs(u)) > 1e-3
    assert lines[1] == "  energy =  0.00e+00     nu =   0.2560"            # nu = D0 dt / (hx hy) = 0.001 * 16 * 16


def test_integrators_agree_to_their_order():
    ref, _, _ = ho.rk3bs(ho.rhs, np.zeros((8, 9)), 1e-4, 0.01, atol=1e-10, rtol=1e-10)
    err = {}
    for name, dts in (("beuler", (1e-3, 5e-4)), ("cn", (1e-3, 5e-4))):
        err[name] = [np.max(np.abs(ho.theta(np.zeros((8, 9)), dt, 0.01, theta=1.0 if name == "beuler" else 0.5)[0] - ref))
                     for dt in dts]
    assert 1.7 < err["beuler"][0] / err["beuler"][1] < 2.3            # first order
    assert 3.4 < err["cn"][0] / err["cn"][1] < 4.6                    # second order
