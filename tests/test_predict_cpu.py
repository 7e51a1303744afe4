"""Host mirror of the reference's motion model (randt::predict / predictSE2 behind Matcher::predictTransform,
R/src/ndt_registration/ndt_matcher.cpp:22-59, R/include/ndt_registration/ceres_residuals.h:25-85): known answers and the two
quirks the reference has — dt is clamped to >= 0.2 s, and the SE(2) model integrates the acceleration into the screw with dt/2
where the vector model uses dt^2/2."""
import math

import numpy as np

from randt_slam_b200 import hostapi


def state(theta=0.0, x=0.0, y=0.0, v=(0.0, 0.0), w=0.0, a=(0.0, 0.0)):
    return np.array([math.cos(theta), math.sin(theta), x, y, x, y, theta, v[0], v[1], w, a[0], a[1]])


def test_vector_model_known_answers():
    s = state(theta=0.3, x=1.0, y=-2.0, v=(2.0, 0.5), w=0.4, a=(0.2, -0.1))
    dt = 0.25
    o = hostapi.predict(s, dt, se2_model=False)
    rot_mid = 0.3 + 0.5 * dt * 0.4
    dx = 2.0 * dt + 0.5 * 0.2 * dt * dt; dy = 0.5 * dt + 0.5 * -0.1 * dt * dt
    assert abs(o[4] - (1.0 + math.cos(rot_mid) * dx - math.sin(rot_mid) * dy)) < 1e-15
    assert abs(o[5] - (-2.0 + math.sin(rot_mid) * dx + math.cos(rot_mid) * dy)) < 1e-15
    assert abs(o[6] - (0.3 + dt * 0.4)) < 1e-15
    assert abs(o[7] - (2.0 + dt * 0.2)) < 1e-15 and abs(o[8] - (0.5 - dt * 0.1)) < 1e-15 and o[9] == 0.4
    # angles come back normalised to (-pi, pi]
    o2 = hostapi.predict(state(theta=3.1, w=1.0), 0.25, se2_model=False)
    assert -math.pi <= o2[6] < math.pi and abs(o2[6] - (3.1 + 0.25 - 2 * math.pi)) < 1e-12


def test_dt_is_clamped_to_200_ms():
    s = state(theta=0.1, v=(1.0, 0.0), w=0.2)
    for se2 in (False, True):
        a = hostapi.predict(s, 0.0, se2); b = hostapi.predict(s, 0.2, se2); c = hostapi.predict(s, -1.0, se2)
        assert np.array_equal(a, b) and np.array_equal(a, c)


def test_se2_model_is_the_exponential_of_the_screw():
    theta, x, y, v, w, a = 0.4, 3.0, 1.0, (1.5, -0.3), 0.6, (0.4, 0.2)
    dt = 0.3
    o = hostapi.predict(state(theta, x, y, v, w, a), dt, se2_model=True)
    ux, uy, th = v[0] * dt + 0.5 * dt * a[0], v[1] * dt + 0.5 * dt * a[1], w * dt       # dt / 2, not dt^2 / 2 (ceres_residuals.h:77-79)
    sbt, omc = math.sin(th) / th, (1 - math.cos(th)) / th
    ex, ey = sbt * ux - omc * uy, omc * ux + sbt * uy
    assert abs(o[2] - (x + math.cos(theta) * ex - math.sin(theta) * ey)) < 1e-14
    assert abs(o[3] - (y + math.sin(theta) * ex + math.cos(theta) * ey)) < 1e-14
    assert abs(math.atan2(o[1], o[0]) - (theta + th)) < 1e-14 and abs(math.hypot(o[0], o[1]) - 1.0) < 1e-15
    # zero turn rate: the small-angle series, a straight step along the heading
    o0 = hostapi.predict(state(theta, x, y, (2.0, 0.0), 0.0), 0.25, se2_model=True)
    assert abs(o0[2] - (x + 0.5 * math.cos(theta))) < 1e-15 and abs(o0[3] - (y + 0.5 * math.sin(theta))) < 1e-15


def test_models_agree_for_constant_velocity_without_rotation():
    s = state(theta=-0.7, x=0.5, y=0.25, v=(1.2, 0.4))
    a = hostapi.predict(s, 0.25, False); b = hostapi.predict(s, 0.25, True)
    assert abs(a[4] - b[2]) < 1e-15 and abs(a[5] - b[3]) < 1e-15
