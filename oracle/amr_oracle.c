/* amr_oracle.c -- TEST INFRASTRUCTURE ONLY (#included by mhd_oracle.c).
 * CPU restatement of the reference's refinement criteria (dynamic refinement, first part):
 *   src/mesh/refinement_criteria.cpp:177-238 (CheckMinMax), :243-287 (CheckSlope), :292-336 (CheckSecondDeriv),
 *   :341-386 (CheckLocation).  q = one variable of a cell-centred array (e.g. w0 with var = IDN for mhd_w_d);
 * flag[nmb] is updated in place exactly as the reference updates refine_flag (a block already flagged for
 * refinement by an earlier criterion is never turned into a derefinement).
 * Parity: tests/test_amr_cpu.py follows an unmodified reference AMR run dump by dump. */
enum { AMR_MIN_MAX = 0, AMR_SLOPE = 1, AMR_SECOND_DERIV = 2, AMR_LOCATION = 3 };

void oracle_amr_criterion(const ocfg *c, int method, const double *a, int nv, int var, double value_min, double value_max,
                          const double *loc4 /* x1, x2, x3, rad */, const double *mb_size, int *flag) {
  odims D = dims(c);
  const int multi_d = c->nx2 > 1, three_d = c->nx3 > 1;
  for (int m = 0; m < c->nmb; ++m) {
    if (method == AMR_LOCATION) {
      const double *sz = mb_size + (size_t)m * 9;
      const double x1 = loc4[0], x2 = loc4[1], x3 = loc4[2], rad = loc4[3];
      const double x1min = sz[0], x2min = sz[1], x3min = sz[2], x1max = sz[3], x2max = sz[4], x3max = sz[5];
      const int in1 = ((x1min < (x1 + rad)) && (x1min > (x1 - rad))) || ((x1max < (x1 + rad)) && (x1max > (x1 - rad))) || ((x1max > (x1 + rad)) && (x1min < (x1 - rad)));
      const int in2 = ((x2min < (x2 + rad)) && (x2min > (x2 - rad))) || ((x2max < (x2 + rad)) && (x2max > (x2 - rad))) || ((x2max > (x2 + rad)) && (x2min < (x2 - rad)));
      const int in3 = ((x3min < (x3 + rad)) && (x3min > (x3 - rad))) || ((x3max < (x3 + rad)) && (x3max > (x3 - rad))) || ((x3max > (x3 + rad)) && (x3min < (x3 - rad)));
      if (in1 && (!multi_d || in2) && (!three_d || in3)) flag[m] = 1;
      continue;
    }
    double qmax = -FLT_MAX, qmin = FLT_MAX;      /* Kokkos::Max / Kokkos::Min start from their identities */
    for (int k = D.ks; k <= D.ke; ++k)
      for (int j = D.js; j <= D.je; ++j)
        for (int i = D.is; i <= D.ie; ++i) {
          const double q = CC(a, nv, m, var, k, j, i);
          double v;
          if (method == AMR_MIN_MAX) { qmin = fmin(q, qmin); v = q; }
          else if (method == AMR_SLOPE) {
            double d2 = SQR(CC(a, nv, m, var, k, j, i + 1) - CC(a, nv, m, var, k, j, i - 1));
            if (multi_d) d2 += SQR(CC(a, nv, m, var, k, j + 1, i) - CC(a, nv, m, var, k, j - 1, i));
            if (three_d) d2 += SQR(CC(a, nv, m, var, k + 1, j, i) - CC(a, nv, m, var, k - 1, j, i));
            v = 0.5 * sqrt(d2) / q;
          } else {
            double d2q = CC(a, nv, m, var, k, j, i + 1) - 2.0 * q + CC(a, nv, m, var, k, j, i - 1);
            if (multi_d) d2q += (CC(a, nv, m, var, k, j + 1, i) - 2.0 * q + CC(a, nv, m, var, k, j - 1, i));
            if (three_d) d2q += (CC(a, nv, m, var, k + 1, j, i) - 2.0 * q + CC(a, nv, m, var, k - 1, j, i));
            v = fabs(d2q) / q;
          }
          qmax = fmax(v, qmax);
        }
    if (value_max < FLT_MAX) {
      if (qmax > value_max) flag[m] = 1;
      if ((qmax < value_max) && (flag[m] == 0)) flag[m] = -1;
    }
    if (method == AMR_MIN_MAX && value_min > -FLT_MAX) {
      if (qmin < value_min) flag[m] = 1;
      if ((qmin > value_min) && (flag[m] == 0)) flag[m] = -1;
    }
  }
}
