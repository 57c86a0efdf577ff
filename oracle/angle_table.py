"""TEST INFRASTRUCTURE ONLY -- per-angle rotation table, evaluated with NumPy the
way the reference does (reference sea_ice_drift/pmlib.py:105-110), so that the C
oracle never has to call libm's sin/cos."""
import numpy as np


def angle_table(angles, alpha0, img_size):
    """Return (n_angles, 4) float64: cos a, sin a, tcdot0, tcdot1 with
    a = radians(angle - alpha0) and tcdot = [tc, tc] . [[cos,-sin],[sin,cos]]."""
    tc = int(img_size / 2.) + 1
    tcv = np.array([tc, tc])
    tab = np.zeros((len(angles), 4), dtype=np.float64)
    for k, angle in enumerate(angles):
        a = np.radians(angle - alpha0)
        transform = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        tcdot = tcv.dot(transform)
        tab[k] = (transform[0, 0], transform[1, 0], tcdot[0], tcdot[1])
    return tab
