// reagent_b200 -- dense feature preprocessing, one descriptor per OUTPUT column.
//
// Restates Preprocessor.forward (reagent/preprocessing/preprocessor.py:115-170 and the
// per-type transforms :202-525).  The reference sorts features by type and splits/cats
// sections; here every output column j carries {source column, type, parameters}, so ENUM
// one-hot expansion and the type sections need no data movement and writes are coalesced.
#pragma once
#include "rb200_common.cuh"

namespace rb200 {

// reagent/preprocessing/normalization.py:34-36
#define RB200_MAX_FEATURE_VALUE 11.513f
#define RB200_MIN_FEATURE_VALUE (-11.513f)

__device__ __forceinline__ float preprocess_value(float x, float presence,
                                                  const rb200_feature_col_t& f,
                                                  const float* __restrict__ quantiles) {
  float y;
  switch (f.type) {
    case RB200_FT_BINARY:  // preprocessor.py:210-217
      y = 1.f - ((x == 0.f) ? 1.f : 0.f);
      break;
    case RB200_FT_PROBABILITY: {  // :237-246
      const float c = fminf(fmaxf(x, f.p0), f.p1);
      y = -1.f * logf((1.f / c) - 1.f);
      break;
    }
    case RB200_FT_CONTINUOUS:  // :315-324
      y = (x - f.p0) / f.p1;
      break;
    case RB200_FT_BOXCOX: {  // :346-364 ; p0=mean p1=std p2=shift p3=lambda
      const float b = (powf(fmaxf(x + f.p2, 1e-6f), f.p3) - 1.f) / f.p3;
      y = (b - f.p0) / f.p1;
      break;
    }
    case RB200_FT_ENUM:  // :518-525 ; p0 = the possible value of this output column
      y = (x == f.p0) ? 1.f : 0.f;
      break;
    case RB200_FT_QUANTILE: {  // :434-505 ; p0=num_quantiles p1=max p2=min
      const float* qb = quantiles + f.q_off;
      float left = -1e20f, right = 1e20f, ge_count = 0.f;
      for (int i = 0; i < f.q_cnt; ++i) {
        const float b = qb[i];
        if (x >= b) { ge_count += 1.f; left = fmaxf(left, b); }
        else { right = fminf(right, b); }
      }
      const float set_to_max = (x >= f.p1) ? 1.f : 0.f;
      const float set_to_min = (x <= f.p2) ? 1.f : 0.f;
      const float interpolate = ((set_to_min + set_to_max) < 0.01f) ? 1.f : 0.f;
      const float left_start = ge_count - 1.f;
      const float interp = (left_start + ((x - left) / ((right + 1e-6f) - left))) / f.p0;
      y = set_to_max + (interpolate * interp);
      break;
    }
    case RB200_FT_CONTINUOUS_ACTION: {  // :274-286 ; p0=min_serving p1=scale p2=min_training
      const float c = (x - f.p0) * f.p1 + f.p2;
      y = fminf(fmaxf(c, f.p3), -f.p3);  // p3 = -1+EPS, -p3 = 1-EPS
      break;
    }
    case RB200_FT_CLIP_LOG:  // :224-230
      y = logf(fmaxf(x, 1e-6f));
      break;
    default:  // DISCRETE_ACTION / DO_NOT_PREPROCESS: identity
      y = x;
      break;
  }
  y = y * presence;
  if (f.type != RB200_FT_DO_NOT_PREPROCESS)  // preprocessor.py:164-167
    y = fminf(fmaxf(y, RB200_MIN_FEATURE_VALUE), RB200_MAX_FEATURE_VALUE);
  return y;
}

}  // namespace rb200
