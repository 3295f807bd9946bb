// smr_fc_emul.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the per-element bodies of the face-centred SMR kernels
// (athenak_b200/csrc/smr_fc_body.h, the same source the CUDA kernels wrap) in host loops over the tables that
// ak_smr_create built, so tests/test_smr_fc_emul_cpu.py can compare the index logic and the arithmetic with
// oracle/smr_fc_oracle.c without a GPU.  Compiled with g++ by the test; never part of the library.
#include "../../athenak_b200/csrc/smr_fc_body.h"

using namespace akb_smr;

extern "C" {

void emul_restrict_fc(const FcDims *D, const double *b1, const double *b2, const double *b3, double *cb1, double *cb2, double *cb3) {
  const double *b[3] = {b1, b2, b3};
  double *cb[3] = {cb1, cb2, cb3};
  for (int v = 0; v < 3; ++v)
    for (long t = 0; t < fc_restrict_count(*D, v); ++t) fc_restrict_elem(*D, v, t, b[v], cb[v]);
}

// reverse = 1 walks the elements backwards: the result of a gather must not depend on the order
void emul_exchange_fc(const FcDims *D, const unsigned char *const *maps /* f0,f1,f2,c0,c1,c2 */, const FcSrc *src,
                      double *b1, double *b2, double *b3, double *cb1, double *cb2, double *cb3, int reverse) {
  double *b[3] = {b1, b2, b3}, *cb[3] = {cb1, cb2, cb3};
  for (int seg = 0; seg < 6; ++seg) {
    const int s = reverse ? 5 - seg : seg, v = s % 3, coarse = s / 3;
    const int per = (int)(coarse ? D->cf[v] : D->f[v]);
    for (int mm = 0; mm < D->nmb; ++mm)
      for (int q = 0; q < per; ++q)
        fc_exchange_elem(*D, v, coarse, reverse ? D->nmb - 1 - mm : mm, reverse ? per - 1 - q : q, maps[s], src, b[v], cb[v]);
  }
}

void emul_prolongate_fc(const FcDims *D, const FcBox *fill, long nfill, const FcBox *shared, long nshared, const FcCell *interior,
                        long ninterior, int steps, double *b1, double *b2, double *b3, double *cb1, double *cb2, double *cb3, int reverse) {
  if (steps & 1)
    for (long q = 0; q < nfill; ++q) {
      const FcBox &d = fill[reverse ? nfill - 1 - q : q];
      for (long t = 0; t < (long)d.n[0] * d.n[1] * d.n[2]; ++t) fc_fill_elem(*D, d, t, b1, b2, b3, cb1, cb2, cb3);
    }
  if (steps & 2) {
    for (long q = 0; q < nshared; ++q) {
      const FcBox &d = shared[reverse ? nshared - 1 - q : q];
      for (long t = 0; t < (long)d.n[0] * d.n[1] * d.n[2]; ++t) fc_prolong_shared_elem(*D, d, t, b1, b2, b3, cb1, cb2, cb3);
    }
    for (long q = 0; q < ninterior; ++q) {
      const FcCell &d = interior[reverse ? ninterior - 1 - q : q];
      for (long t = 0; t < (long)d.n[0] * d.n[1] * d.n[2]; ++t) fc_prolong_interior_elem(*D, d, t, b1, b2, b3);
    }
  }
}

void emul_emf_correct(const FcDims *D, const EmfPack *pack, long npack, const EmfOp *ops, const EmfRange *ranges, double *buf,
                      double *e1, double *e2, double *e3, int reverse) {
  for (long q = 0; q < npack; ++q) {
    const EmfPack &d = pack[q];
    for (long t = 0; t < (long)d.n[0] * d.n[1] * d.n[2]; ++t) emf_pack_elem(*D, d, t, e1, e2, e3, buf);
  }
  double *e[3] = {e1, e2, e3};
  for (int v = 0; v < 3; ++v) {
    const int ns = emf_surface_count(*D, v);
    for (int m = 0; m < D->nmb; ++m)
      for (int q = 0; q < ns; ++q) emf_apply_elem(*D, v, reverse ? D->nmb - 1 - m : m, reverse ? ns - 1 - q : q, ranges, ops, buf, e[v]);
  }
}

}  // extern "C"
