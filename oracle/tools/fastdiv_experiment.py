#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- reproduces on the CPU the one bitwise difference the GPU showed in a long AMR run
(DESIGN.md 6b).  Builds a copy of the oracle whose PLM limiter divides with the DEVICE's guard-free formula
(reciprocal, q = a*r, one remainder step: athenak_b200/csrc/mhd_device.cuh ddiv) instead of IEEE division, optionally with
the guard the device PLM now has (numerator below 2^-900 and non-zero -> IEEE), and follows the reference's 2-D three-level
AMR blast cycle by cycle (tests/amr_util.py).  Needs oracle/_ref/athena_blast.

  python oracle/tools/fastdiv_experiment.py            # guard-free: differs in cycle 24 only, 4 faces of x2f, 5e-175
  python oracle/tools/fastdiv_experiment.py --guard    # guarded: all cycles bitwise
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


def main():
    guard = "--guard" in sys.argv
    tmp = tempfile.mkdtemp(prefix="fastdiv_")
    try:
        for f in os.listdir(os.path.join(ROOT, "oracle")):
            if f.endswith(".c"):
                shutil.copy(os.path.join(ROOT, "oracle", f), tmp)
        src = open(os.path.join(tmp, "mhd_oracle.c")).read()
        assert OLD in src
        open(os.path.join(tmp, "mhd_oracle.c"), "w").write(src.replace(OLD, NEW, 1))
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
        print(f"guard={guard}: {nbad} of {len(dumps) - 2} cycles differ from the reference")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
