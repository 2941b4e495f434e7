// rb_offline.cu -- offline basis linear algebra (Gram S^T X S2, Galerkin projections).  Placeholder until the kernels
// land: returns RB_ERR_UNSUPPORTED (fails loudly; no CPU fallback).
#include "rb_common.cuh"

extern "C" int rb_gram(rb_ctx* ctx, int64_t, int, int, const double*, const double*, const int64_t*, const int32_t*,
                       const double*, int64_t, int64_t, double*) {
  if (!ctx) return RB_ERR_ARG;
  RB_FAIL(RB_ERR_UNSUPPORTED, "rb_gram: not implemented yet");
}
extern "C" int rb_project(rb_ctx* ctx, int64_t, int, int, const double*, const int64_t*, const int32_t*, const double*,
                          const int64_t*, double*) {
  if (!ctx) return RB_ERR_ARG;
  RB_FAIL(RB_ERR_UNSUPPORTED, "rb_project: not implemented yet");
}
