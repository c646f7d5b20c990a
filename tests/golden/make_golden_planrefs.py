"""Generate tests/golden/planner_refs.npz: the planner's reference post-processing (plannerMain.py:196-224, 257-280)
restated line by line on top of the REFERENCE'S OWN Map.getGlobalPosition and the same SciPy calls the reference makes
(interp1d(kind='cubic'), signal.ellip(4, 0.01, 120, 0.125), signal.filtfilt(padlen=50)), fed with the plans recorded in
tests/golden/planner_loop.npz.

    python tests/golden/make_golden_planrefs.py
"""
import os
import sys

import numpy as np
from scipy import signal
from scipy.interpolate import interp1d

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ns = refload.load()
    m = ns.Map()
    G = np.load(os.path.join(OUT, "planner_loop.npz"))
    b_filter, a_filter = signal.ellip(4, 0.01, 120, 0.125)   # plannerMain.py:112
    dt = 1.0 / 20.0
    xin, ssin, x0in, outs = [], [], [], []
    for case in (0, 1):
        XP, SSs = G["loop%d_xpred" % case], G["loop%d_SS" % case]
        N = XP.shape[1] - 1
        Xlast = Ylast = Thetalast = 0.0
        for t in range(XP.shape[0]):
            xPred, SS = XP[t], SSs[t]
            Xref, Yref, Thetaref = np.zeros(N + 1), np.zeros(N + 1), np.zeros(N + 1)
            Xref[0], Yref[0], Thetaref[0] = Xlast, Ylast, Thetalast            # :196-198
            for j in range(0, N):                                              # :203-209 (SS[j+1] already integrated)
                Xref[j + 1], Yref[j + 1], Thetaref[j + 1] = m.getGlobalPosition(SS[j + 1], 0.0)
            x0in.append([Xlast, Ylast, Thetalast])
            Xlast, Ylast, Thetalast = Xref[1], Yref[1], Thetaref[1]             # :214-216
            xp, yp, yaw = np.zeros(N), np.zeros(N), np.zeros(N)
            for i in range(0, N):                                              # :218-221
                yaw[i] = Thetaref[i] + xPred[i, 4]
                xp[i] = Xref[i] - xPred[i, 3] * np.sin(yaw[i])
                yp[i] = Yref[i] + xPred[i, 3] * np.cos(yaw[i])
            vel = xPred[0:N, 0]                                                # :223-224
            curv = xPred[0:N, 2] / xPred[0:N, 0]
            interp_dt = 0.033                                                  # :257-280
            time50ms = np.linspace(0, N * dt, num=N, endpoint=True)
            time33ms = np.linspace(0, N * dt, num=int(np.around(N * dt / interp_dt)), endpoint=True)
            res = [interp1d(time50ms, v, kind='cubic')(time33ms) for v in (xp, yp, yaw, vel, curv)]
            res[4] = signal.filtfilt(b_filter, a_filter, res[4], padlen=50)
            xin.append(xPred); ssin.append(SS); outs.append(np.array(res))
    np.savez_compressed(os.path.join(OUT, "planner_refs.npz"), x_pred=np.array(xin), SS=np.array(ssin), xyth0=np.array(x0in),
                        refs=np.array(outs), ellip_b=b_filter, ellip_a=a_filter)
    print("planner_refs.npz", np.array(outs).shape, os.path.getsize(os.path.join(OUT, "planner_refs.npz")), "bytes")


if __name__ == "__main__":
    main()
