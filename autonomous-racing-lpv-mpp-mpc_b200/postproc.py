"""Planner -> controller reference post-processing (plannerMain.py:212-224, 257-280, 112).

For a uniform input grid the reference's ``interp1d(time50ms, y, kind='cubic')(time33ms)`` (not-a-knot cubic spline) and
``signal.filtfilt(b, a, y, padlen=50)`` (odd extension, ``lfilter_zi`` initial conditions, forward-backward IIR) are
LINEAR maps of the samples, so both are built here once as dense fp64 matrices and the device applies them to every plan
of a batch (``lpvmpc_plan_refs_*``): ``W`` [n_out, n_in] for x, y, yaw, vx and ``F W`` for the curvature.

No SciPy on the product path: the spline and the filter are written out below; tests compare them with SciPy, which is
what the reference calls.
"""
import numpy as np

# signal.ellip(4, 0.01, 120, 0.125) (plannerMain.py:112): 4th-order elliptic low-pass, 0.01 dB ripple, 120 dB stop band,
# corner 0.125 of Nyquist.  Constants of the reference's configuration (checked against SciPy in the tests).
ELLIP_B = np.array([0.002371156562865973, 0.009078139816074575, 0.013422870472257976, 0.009078139816074575, 0.0023711565628659736])
ELLIP_A = np.array([1.0, -2.7936694495096583, 3.135346099277687, -1.6401318189059317, 0.33481847307876444])
INTERP_DT = 0.033   # plannerMain.py:257
PADLEN = 50         # plannerMain.py:280


def n_out_for(N, dt, interp_dt=INTERP_DT):
    """np.around(N*dt/interp_dt) of plannerMain.py:259."""
    return int(np.around(N * dt / interp_dt))


def spline_matrix(n_in, n_out, T):
    """W with W @ y == interp1d(linspace(0,T,n_in), y, kind='cubic')(linspace(0,T,n_out)): not-a-knot cubic spline."""
    x = np.linspace(0.0, T, num=n_in, endpoint=True)
    t = np.linspace(0.0, T, num=n_out, endpoint=True)
    h = x[1] - x[0]
    # second derivatives M = S @ y: interior continuity rows + not-a-knot end rows (uniform grid)
    A = np.zeros((n_in, n_in))
    R = np.zeros((n_in, n_in))
    for i in range(1, n_in - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = h, 4.0 * h, h
        R[i, i - 1], R[i, i], R[i, i + 1] = 6.0 / h, -12.0 / h, 6.0 / h
    A[0, 0], A[0, 1], A[0, 2] = 1.0, -2.0, 1.0
    A[-1, -1], A[-1, -2], A[-1, -3] = 1.0, -2.0, 1.0
    S = np.linalg.solve(A, R)
    W = np.zeros((n_out, n_in))
    for r, tt in enumerate(t):
        i = min(int(np.floor(tt / h + 1e-12)), n_in - 2)
        a, b = x[i + 1] - tt, tt - x[i]
        W[r] = S[i] * (a ** 3 / (6 * h) - h * a / 6) + S[i + 1] * (b ** 3 / (6 * h) - h * b / 6)
        W[r, i] += a / h
        W[r, i + 1] += b / h
    return W


def _lfilter(b, a, x, zi):
    """Direct-form II transposed IIR along axis 0 of x [n, m] with initial state zi [order, m]."""
    n = x.shape[0]
    z = zi.copy()
    y = np.empty_like(x)
    order = len(a) - 1
    for i in range(n):
        y[i] = b[0] * x[i] + z[0]
        for k in range(order - 1):
            z[k] = b[k + 1] * x[i] + z[k + 1] - a[k + 1] * y[i]
        z[order - 1] = b[order] * x[i] - a[order] * y[i]
    return y


def _lfilter_zi(b, a):
    order = len(a) - 1
    comp = np.zeros((order, order))
    comp[0, :] = -a[1:]
    comp[1:, :-1] = np.eye(order - 1)
    return np.linalg.solve(np.eye(order) - comp.T, b[1:] - a[1:] * b[0])


def filtfilt_matrix(b, a, n, padlen=PADLEN):
    """F with F @ y == signal.filtfilt(b, a, y, padlen=padlen) (padtype 'odd', method 'pad') for len(y) == n."""
    if padlen >= n:
        raise ValueError("padlen must be smaller than the signal length")
    X = np.eye(n)
    left = 2 * X[0:1] - X[padlen:0:-1]
    right = 2 * X[-1:] - X[-2:-(padlen + 2):-1]
    ext = np.concatenate([left, X, right], axis=0)
    zi = _lfilter_zi(b, a)
    y = _lfilter(b, a, ext, zi[:, None] * ext[0][None, :])
    yr = y[::-1]
    y2 = _lfilter(b, a, yr, zi[:, None] * yr[0][None, :])
    return y2[::-1][padlen:-padlen]


def reference_matrices(N, dt):
    """(W, Wc): x / y / yaw / vx resampling and curvature resampling + filtering for horizon N and planner period dt."""
    n_out = n_out_for(N, dt)
    W = spline_matrix(N, n_out, N * dt)
    F = filtfilt_matrix(ELLIP_B, ELLIP_A, n_out)
    return np.ascontiguousarray(W), np.ascontiguousarray(F @ W)
