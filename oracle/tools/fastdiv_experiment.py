#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- reproduces on the CPU the one bitwise difference the GPU showed in a long AMR run
(DESIGN.md 6b).  Builds a copy of the oracle whose PLM limiter divides with the DEVICE's guard-free formula
(reciprocal, q = a*r, one remainder step: athenak_b200/csrc/mhd_device.cuh ddiv) instead of IEEE division, optionally with
the guard the device PLM now has (numerator below 2^-900 and non-zero -> IEEE), and follows the reference's 2-D three-level
AMR blast cycle by cycle (tests/amr_util.py).  Needs oracle/_ref/athena_blast.

  python oracle/tools/fastdiv_experiment.py            # guard-free: differs in cycle 24 only, 4 faces of x2f, 5e-175
  python oracle/tools/fastdiv_experiment.py --guard    # guarded: all cycles bitwise
  python oracle/tools/fastdiv_experiment.py --ppm4     # the headline path's div_by (Markstein x/12, x/dx in PPM4, the update and
                                                       # CT) instead: PPM4+HLLD+RK3, 69 cycles, no cycle differs (the PPM4
                                                       # precursor stays above 1e-60, far from the denormal floor)
"""
import ctypes as C
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

OLD = "  double dqm = dq2 / (dql + dqr);\n  if (dq2 <= 0.0) dqm = 0.0;"
NEW = """  double dqm;
  {
    const double a = dq2, b = (dql + dqr);
    if (GUARD && fabs(a) < 0x1p-900 && a != 0.0) dqm = a / b;
    else {
      const double r = 1.0 / b;            /* the correctly rounded reciprocal stands in for the converged Newton iterate */
      const double qq = a * r;
      const double rem = fma(qq, -b, a);
      dqm = fma(r, rem, qq);
    }
  }
  if (dq2 <= 0.0) dqm = 0.0;"""


DIVBY = """
static double DIVBY(double x, double y) { const double r = 1.0 / y; const double q = x * r; const double e = fma(-y, q, x); return fma(e, r, q); }
static void PPM4("""
PPM4_SUBS = [
    ("\nstatic void PPM4(", DIVBY),
    ("double qlv = (7. * (q + qm1) - (qm2 + qp1)) / 12.0;\n  double qrv = (7. * (q + qp1) - (qm1 + qp2)) / 12.0;\n  qlv = fmax(qlv, fmin(q, qm1));",
     "double qlv = DIVBY(7. * (q + qm1) - (qm2 + qp1), 12.0);\n  double qrv = DIVBY(7. * (q + qp1) - (qm1 + qp2), 12.0);\n  qlv = fmax(qlv, fmin(q, qm1));"),
    ("(F1(flx1, m, n, k, j, i + 1) - F1(flx1, m, n, k, j, i)) / dx1", "DIVBY(F1(flx1, m, n, k, j, i + 1) - F1(flx1, m, n, k, j, i), dx1)"),
    ("(F2(flx2, m, n, k, j + 1, i) - F2(flx2, m, n, k, j, i)) / dx2", "DIVBY(F2(flx2, m, n, k, j + 1, i) - F2(flx2, m, n, k, j, i), dx2)"),
    ("b -= beta_dt * (X3E(e3, m, k, j + 1, i) - X3E(e3, m, k, j, i)) / dx2;", "b -= DIVBY(beta_dt * (X3E(e3, m, k, j + 1, i) - X3E(e3, m, k, j, i)), dx2);"),
    ("b += beta_dt * (X3E(e3, m, k, j, i + 1) - X3E(e3, m, k, j, i)) / dx1;", "b += DIVBY(beta_dt * (X3E(e3, m, k, j, i + 1) - X3E(e3, m, k, j, i)), dx1);"),
]


def main():
    guard = "--guard" in sys.argv
    ppm4 = "--ppm4" in sys.argv
    tmp = tempfile.mkdtemp(prefix="fastdiv_")
    try:
        for f in os.listdir(os.path.join(ROOT, "oracle")):
            if f.endswith(".c"):
                shutil.copy(os.path.join(ROOT, "oracle", f), tmp)
        src = open(os.path.join(tmp, "mhd_oracle.c")).read()
        if ppm4:
            for a, b in PPM4_SUBS:
                assert a in src, a
                src = src.replace(a, b, 1)
        else:
            assert OLD in src
            src = src.replace(OLD, NEW, 1)
        open(os.path.join(tmp, "mhd_oracle.c"), "w").write(src)
        so = os.path.join(tmp, "liboracle_exp.so")
        subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-fPIC", "-ffp-contract=off", "-fno-fast-math", f"-DGUARD={int(guard)}", "-shared",
                               "-o", so, os.path.join(tmp, "mhd_oracle.c"), "-lm"])
        import numpy as np
        import oracle_lib as O
        O._lib = C.CDLL(so)
        for n in ("oracle_new_dt", "oracle_initial_dt"):
            getattr(O._lib, n).restype = C.c_double
        import amr_util as AU
        ov = ("mesh_refinement/num_levels=3", "time/nlim=36")
        if ppm4:
            ov = ("mesh/nghost=4", "mhd/reconstruct=ppm4", "time/integrator=rk3", "time/nlim=70", "mesh/nx1=64", "mesh/nx2=64")
        dumps = AU.reference_dumps(ov)
        S = AU.Stepper(AU.params(ov))
        nbad = 0
        for k in range(1, len(dumps) - 1):
            r, rn = dumps[k], dumps[k + 1]
            A, t, dt = S.advance(r, k)
            du = int(np.count_nonzero(A.u0 != rn.u0))
            db = [int(np.count_nonzero(x != y)) for x, y in zip(A.b0, (rn.b1f, rn.b2f, rn.b3f))]
            if du or any(db) or t != rn.time or dt != rn.dt:
                nbad += 1
                mx = [float(np.abs(x - y).max()) for x, y in zip(A.b0, (rn.b1f, rn.b2f, rn.b3f))]
                print(f"cycle {k}: U faces differing {du}, B faces differing {db}, max|dB| {mx}, time equal {t == rn.time}, dt equal {dt == rn.dt}")
        print(f"{'ppm4 div_by' if ppm4 else 'plm ddiv'} guard={guard}: {nbad} of {len(dumps) - 2} cycles differ from the reference")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
