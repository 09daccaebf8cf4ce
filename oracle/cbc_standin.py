#!/usr/bin/env python3
"""Stand-in for the external `cbc` binary the reference shells out to (localhap.cpp:179-181).

TEST INFRASTRUCTURE (this container only): COIN-OR Cbc 2.10.8 (environment.yml:14,19) is not
installed and cannot be fetched, so the reference binary built by `make -C oracle ref` calls this
script through a `cbc` wrapper placed first on PATH.  Usage, as the reference invokes it:

    cbc <prefix>.lp solve solu <prefix>.sol

It reads the "AMBI-LP 1" dump that the COIN shim's writeLp() produced from the reference's own
model-building code, and writes a Cbc-style solution file (the reference parses the word after
"value" as the objective and every `x<idx> <value>` pair, localhap.cpp:194-212).

Modes (env AMBI_CBC_MODE):
  highs   (default) solve the dumped MILP exactly with HiGHS (scipy.optimize.milp)
  inject  do not solve: take the integer assignment from the file named by AMBI_CBC_SOL
          (one `idx value` per line) -- used to push OUR Stage-A answer through the reference's
          own Stage B and compare the printed path.
"""
import os
import sys

import numpy as np


def read_lp(fn):
    with open(fn) as f:
        assert f.readline().strip() == "AMBI-LP 1"
        nrows, ncols = map(int, f.readline().split())
        rows, rlo, rhi = [], [], []
        for _ in range(nrows):
            t = f.readline().split()
            assert t[0] == "R"
            rlo.append(float(t[1])); rhi.append(float(t[2]))
            rows.append([(int(p.split(":")[0]), float(p.split(":")[1])) for p in t[4:]])
        clo, chi, obj, integ = [], [], [], []
        for _ in range(ncols):
            t = f.readline().split()
            assert t[0] == "C"
            clo.append(float(t[1])); chi.append(float(t[2])); obj.append(float(t[3])); integ.append(int(t[4]))
    return rows, np.array(rlo), np.array(rhi), np.array(clo), np.array(chi), np.array(obj), np.array(integ)


def main():
    lp, sol = sys.argv[1], sys.argv[4]
    rows, rlo, rhi, clo, chi, obj, integ = read_lp(lp)
    ncols = len(obj)
    mode = os.environ.get("AMBI_CBC_MODE", "highs")
    if mode == "inject":
        x = np.zeros(ncols)
        nint = int(integ.sum())
        for line in open(os.environ["AMBI_CBC_SOL"]):
            i, v = line.split()
            x[int(i)] = float(v)
        # eps columns follow the integer block; the last column is the fixed bias (LGM:4409-4410, 4729-4730)
        x[ncols - 1] = clo[ncols - 1]
        for k in range(nint, ncols - 1):
            r = 2 * (k - nint)
            val = sum(c * x[i] for i, c in rows[r] if i < nint)
            x[k] = abs(val - rlo[r])
        status, fun = 0, float(obj @ x)
    else:
        from scipy.optimize import Bounds, LinearConstraint, milp
        from scipy.sparse import lil_matrix
        a = lil_matrix((len(rows), ncols))
        for r, row in enumerate(rows):
            for i, c in row:
                a[r, i] += c
        res = milp(obj, constraints=LinearConstraint(a.tocsr(), rlo, rhi), integrality=integ, bounds=Bounds(clo, chi),
                   options={"mip_rel_gap": 0.0, "time_limit": float(os.environ.get("AMBI_CBC_TIME", "3600"))})
        status = res.status
        if res.x is None:
            with open(sol, "w") as f:
                f.write("Infeasible - objective value 0.00000000\n")
            return
        x, fun = res.x, float(res.fun)
    with open(sol, "w") as f:
        f.write("Optimal - objective value %.8f\n" % fun)
        for i in range(ncols):
            v = x[i]
            if integ[i]:
                v = round(v)
            if abs(v) > 1e-9:
                f.write("%7d x%-8d %15.8g %15.8g\n" % (i, i, v, 0.0))


if __name__ == "__main__":
    main()
