"""Recover the reference's own closed-loop trajectories from the vector PDFs of its report.

Run in the BUILD container only (reads /root/reference; writes tests/golden/*.npz):

    python tests/golden/make_pdf_golden.py

`PlotUtils.plot_signals` (`/root/reference/HumanoidNavigation/Utils/PlotsUtils.py:21-53`) saved each signal of
`run_simulation_circles` / `run_simulation_circles_custom_ldcbf` (`report_simulations/simulation_1.py:80-192`)
as `Assets/ReportResults/Simulation1Circles{,Delta}/evolutions/evolution_{i}.pdf`.  matplotlib's PDF backend
writes the polyline and the tick marks in points with 6 decimals, so the data are recoverable to ~1e-7:
fit label = a * coordinate + b on the ticks of each axis, apply to the polyline.

evolution_0: position error (p - goal), evolution_1: local (longitudinal, lateral) velocity = R(theta)^T v,
evolution_2: theta, evolution_3: omega (simulation_1.py:108-118).
Output arrays: X[5, K+1] (p_x, v_x, p_y, v_y, theta), omega[K], goal[2], delta.
"""
import os
import re
import sys
import zlib

import numpy as np

REF = "/root/reference/Assets/ReportResults"
NUM = r"-?\d+(?:\.\d+)?"


def content_stream(path):
    data = open(path, "rb").read()
    best = b""
    for s in re.findall(rb"stream\r?\n(.*?)\r?\nendstream", data, re.S):
        try:
            t = zlib.decompress(s)
        except zlib.error:
            continue
        if len(t) > len(best):
            best = t
    return best.decode("latin1")


def parse(path):
    """-> list of polylines (each (n,2) array in data units)."""
    txt = content_stream(path)
    # --- ticks: a 3.5pt stub followed by its label
    xt, yt = [], []
    tick_re = re.compile(rf"({NUM}) ({NUM}) m\n({NUM}) ({NUM}) l\n\nB\n(.*?)(?=\nQ q |\Z)", re.S)
    for m in tick_re.finditer(txt):
        x0, y0, x1, y1 = (float(m.group(i)) for i in range(1, 5))
        body = m.group(5)
        tj = re.search(r"\[(.*?)\] TJ", body, re.S)
        if not tj:
            continue
        label = "".join(re.findall(r"\((.*?)\)", tj.group(1)))
        try:
            val = float(label)
        except ValueError:
            continue
        if "minus Do" in body:
            val = -val
        if abs(x0 - x1) < 1e-9 and abs((y0 - y1) - 3.5) < 1e-6:
            xt.append((x0, val))
        elif abs(y0 - y1) < 1e-9 and abs((x0 - x1) - 3.5) < 1e-6:
            yt.append((y0, val))
    xt, yt = np.array(xt), np.array(yt)
    ax, bx = np.polyfit(xt[:, 0], xt[:, 1], 1)
    ay, by = np.polyfit(yt[:, 0], yt[:, 1], 1)
    resid = max(np.max(np.abs(ax * xt[:, 0] + bx - xt[:, 1])), np.max(np.abs(ay * yt[:, 0] + by - yt[:, 1])))
    # --- polylines: runs of "x y m / x y l ... S" with > 5 points
    lines = []
    for m in re.finditer(rf"((?:{NUM} {NUM} [ml]\n)+)\nS", txt):
        pts = np.array([[float(a), float(b)] for a, b in re.findall(rf"({NUM}) ({NUM}) [ml]", m.group(1))])
        if len(pts) > 5:
            lines.append(np.column_stack((ax * pts[:, 0] + bx, ay * pts[:, 1] + by)))
    return lines, resid


def recover(folder, goal, delta):
    ev = lambda i: parse(f"{REF}/{folder}/evolutions/evolution_{i}.pdf")
    (ex, ey), r0 = ev(0)
    (vl, vt), r1 = ev(1)
    (th,), r2 = ev(2)
    (om,), r3 = ev(3)
    K1 = len(ex)
    assert len(ey) == len(vl) == len(vt) == len(th) == K1 and len(om) == K1 - 1
    theta = th[:, 1]
    px, py = ex[:, 1] + goal[0], ey[:, 1] + goal[1]
    # global v = R(theta) v_local  (PlotUtils.compute_local_velocities rotates by R(theta)^T)
    c, s = np.cos(theta), np.sin(theta)
    vx = c * vl[:, 1] - s * vt[:, 1]
    vy = s * vl[:, 1] + c * vt[:, 1]
    X = np.vstack((px, vx, py, vy, theta))
    t = ex[:, 0]
    return dict(X=X, omega=om[:, 1], t=t, goal=np.array(goal, dtype=float), delta=np.float64(delta),
                tick_fit_residual=np.float64(max(r0, r1, r2, r3)))


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for folder, name, delta in (("Simulation1Circles", "circles_traj", 0.0),
                                ("Simulation1CirclesDelta", "circles_delta_traj", 0.3)):
        g = recover(folder, (6.0, -3.0), delta)
        print(name, g["X"].shape, g["omega"].shape, "tick residual", g["tick_fit_residual"],
              "x0", g["X"][:, 0], "dt", np.diff(g["t"])[:3])
        np.savez(os.path.join(out, name + ".npz"), **g)


if __name__ == "__main__":
    sys.exit(main())
