// amr_emul.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the per-element bodies of the AMR kernels (athenak_b200/csrc/amr_body.h,
// the same source the CUDA kernels wrap) in host loops, so tests/test_amr_emul_cpu.py can compare them with
// oracle/amr_oracle.c without a GPU.  Compiled with g++ by the test; never part of the library.
#include <cfloat>
#include "../../athenak_b200/csrc/amr_body.h"

using namespace akb_smr;

extern "C" {

void emul_amr_move(const FcDims *Dp, int new_nmb, const int *plan, const double *uo, double *un, int nvar,
                   const double *b1o, const double *b2o, const double *b3o, double *b1n, double *b2n, double *b3n) {
  const FcDims &D = *Dp;
  const AmrBlock *blocks = reinterpret_cast<const AmrBlock *>(plan);
  const double *bo[3] = {b1o, b2o, b3o};
  double *bn[3] = {b1n, b2n, b3n};
  for (int m = new_nmb - 1; m >= 0; --m) {          // any order: a gather
    const AmrBlock B = blocks[m];
    for (int v = 0; v < nvar; ++v)
      for (int t = amr_move_cc_count(D) - 1; t >= 0; --t) amr_move_cc_flat(D, B, m, nvar, v, t, uo, un);
    for (int v = 0; v < 3; ++v)
      for (int t = (int)D.f[v] - 1; t >= 0; --t) amr_move_fc_flat(D, B, m, v, t, bo[v], bn[v]);
  }
  for (int m = 0; m < new_nmb; ++m) {
    const AmrBlock B = blocks[m];
    if (B.kind != AMR_REFINE) continue;
    for (int v = 0; v < 3; ++v)
      for (int t = 0; t < amr_refine_fc_shared_count(D, v); ++t) amr_refine_fc_shared_flat(D, B, m, v, t, bo[v], bn[v]);
  }
}

void emul_amr_interior(const FcDims *Dp, int nmb, const int *blockflag, double *b1, double *b2, double *b3) {
  const FcDims &D = *Dp;
  for (int m = 0; m < nmb; ++m) {
    if (!blockflag[m]) continue;
    const FcCell d = amr_whole_block(D, m);
    for (int t = d.n[0] * d.n[1] * d.n[2] - 1; t >= 0; --t) fc_prolong_interior_elem(D, d, t, b1, b2, b3, true);
  }
}

void emul_amr_extrema(const FcDims *Dp, int nmb, int method, int nvar, int var, const double *q, double *out) {
  const FcDims &D = *Dp;
  for (int m = 0; m < nmb; ++m) {
    const double *a = q + ((long)m * nvar + var) * ((long)D.n[0] * D.n[1] * D.n[2]);
    double vmax = -DBL_MAX, vmin = DBL_MAX;
    for (int t = amr_active_count(D) - 1; t >= 0; --t) {
      const double x = amr_criterion_flat(D, method, a, t);
      vmax = fmax(x, vmax); vmin = fmin(x, vmin);
    }
    out[2 * m] = vmax; out[2 * m + 1] = vmin;
  }
}

}  // extern "C"
