// Device math for the ZINB / NB negative log-likelihood and its gradient.
//
// Restates dca/loss.py:72-156 (reference formulas, every epsilon position kept where it is
// representable in fp32) and the closed-form derivatives TF autodiff would produce
// (SURVEY.md A.4), re-arranged so that fp32 does not cancel catastrophically:
//   lgamma(theta)-lgamma(y+theta) = -sum_{k<y} log(theta+k)           (integer y, exact recurrence)
//   psi(theta)-psi(y+theta)       = -sum_{k<y} 1/(theta+k)
//   general (large / non-integer y): shifted Stirling / asymptotic digamma differences
//   d/dtheta nb  = [log1p(x) - x/(1+x)] + y/(theta+mu) - (psi(y+theta)-psi(theta)),  x = mu/theta
//   d/dmu * mu   = theta*(mu-y)/(theta+mu)
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace dca {
namespace zmath {

constexpr float kEps = 1e-10f;            // dca/loss.py:65
constexpr float kHalfLog2Pi = 0.918938533204672742f;

#define DCA_HD __host__ __device__ __forceinline__

// log(k!) for k = 0..15
DCA_HD float log_fact(int k) {
  const float t[16] = {
    0.0f, 0.0f, 0.693147180559945f, 1.791759469228055f, 3.178053830347946f, 4.787491742782046f,
    6.579251212010101f, 8.525161361065415f, 10.60460290274525f, 12.80182748008147f,
    15.10441257307552f, 17.50230784587389f, 19.98721449566189f, 22.55216385312342f,
    25.19122118273868f, 27.89927138384089f};
  return t[k];
}

// lgamma(x) - [(x-0.5)log x - x + 0.5 log 2pi]  for x >= 8
DCA_HD float stirling_corr(float x) {
  float r = 1.0f / x, r2 = r * r;
  return r * (0.0833333333f + r2 * (-0.00277777778f + r2 * 0.000793650794f));
}
// psi(x) - log(x) for x >= 8
DCA_HD float digamma_corr(float x) {
  float r = 1.0f / x, r2 = r * r;
  return -0.5f * r - r2 * (0.0833333333f - r2 * (0.00833333333f - r2 * 0.00396825397f));
}

// prod_{k<8}(x+k) as log, and sum_{k<8} 1/(x+k), for 0 < x < 8
DCA_HD void shift8(float x, float& logprod, float& recsum) {
  float P = 1.f, dP = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { float t = x + (float)k; dP = fmaf(dP, t, P); P *= t; }
  logprod = logf(P); recsum = dP / P;
  P = 1.f; dP = 0.f;
#pragma unroll
  for (int k = 4; k < 8; ++k) { float t = x + (float)k; dP = fmaf(dP, t, P); P *= t; }
  logprod += logf(P); recsum += dP / P;
}

// lgamma(y+1), y >= 0   (only the loss VALUE needs it; it has no gradient)
DCA_HD float lgamma_1p(float y) {
  if (y < 15.5f && y == rintf(y)) return log_fact((int)y);
  float x = y + 1.0f, shift = 0.f;
  if (x < 8.f) { float s; shift8(x, shift, s); x += 8.f; }
  return (x - 0.5f) * logf(x) - x + kHalfLog2Pi + stirling_corr(x) - shift;
}

// lg = lgamma(th+y) - lgamma(th),  dg = psi(th+y) - psi(th);   th > 0, y >= 0
DCA_HD void lgam_digam_diff(float th, float y, float& lg, float& dg) {
  if (y <= 16.f && y == rintf(y)) {
    const int n = (int)y;
    float P = 1.f, dP = 0.f;
    lg = 0.f; dg = 0.f;
    for (int k = 0; k < n; ++k) {
      float t = th + (float)k;
      dP = fmaf(dP, t, P);
      P *= t;
      if ((k & 3) == 3) { lg += logf(P); dg += dP / P; P = 1.f; dP = 0.f; }
    }
    if (n & 3) { lg += logf(P); dg += dP / P; }
    return;
  }
  float a = th, b = th + y, sh_lg = 0.f, sh_dg = 0.f;
  if (a < 8.f) { float lp, rs; shift8(a, lp, rs); sh_lg += lp; sh_dg += rs; a += 8.f; }
  if (b < 8.f) { float lp, rs; shift8(b, lp, rs); sh_lg -= lp; sh_dg -= rs; b += 8.f; }
  const float lr = log1pf((b - a) / a);                 // log(b/a)
  lg = (a - 0.5f) * lr + (b - a) * (logf(b) - 1.0f) + stirling_corr(b) - stirling_corr(a) + sh_lg;
  dg = lr + digamma_corr(b) - digamma_corr(a) + sh_dg;
}

// f(x) = log1p(x) - x/(1+x) >= 0 without cancellation for small x;  L1 = log1p(x) given
DCA_HD float log1p_minus_ratio(float x, float L1) {
  if (x < 0.05f) {
    float p = fmaf(x, -0.857142857f, 0.833333333f);
    p = fmaf(x, p, -0.8f);
    p = fmaf(x, p, 0.75f);
    p = fmaf(x, p, -0.666666667f);
    p = fmaf(x, p, 0.5f);
    return x * x * p;
  }
  return L1 - x / (1.0f + x);
}

struct Elem {
  float loss;   // element NLL (+ ridge*pi^2)
  float gm;     // dL/d zm   (mean pre-activation), clip mask applied, NOT yet / N
  float gd;     // COND_DISP: dL/d zd ; else raw dL/dtheta
  float gp;     // dL/d zp
};

// y: raw count; m: MeanAct output (before *sf); sf: size factor; th: DispAct output or per-gene
// theta; pi: sigmoid output.
template <bool HAS_PI, bool COND_DISP>
DCA_HD Elem zinb_elem(float y, float m, float sf, float th, float pi, float ridge) {
  const float mu = m * sf;                                 // dca/layers.py:85
  const bool m_pass = (m > 1e-5f) && (m < 1e6f);           // clip_by_value gradient mask (network.py:38)
  const float d_in = th;
  th = fminf(th, 1e6f);                                    // dca/loss.py:85
  const float te = th + kEps;
  const float x = mu / te;
  const float L1 = log1pf(x);                              // log(1 + mu/(theta+eps))   loss.py:88
  const float rden = 1.0f / (te + mu);
  const float f = log1p_minus_ratio(x, L1);
  Elem o;
  float dth, dpi = 0.f;
  if (HAS_PI && y < 1e-8f) {                               // loss.py:138  zero branch
    const float z = expf(-th * L1);                        // pow(theta/(theta+mu+eps), theta)  loss.py:136
    const float omp = 1.0f - pi;
    const float D = pi + omp * z + kEps;                   // loss.py:137
    const float rD = 1.0f / D;
    o.loss = -logf(D);
    const float w = omp * z * rD;
    o.gm = w * th * mu * rden;
    dth = w * f;                                           // -w*(log r + 1 - r)
    dpi = -(1.0f - z) * rD;
  } else {                                                 // NB branch  loss.py:87-88,130
    float lg, dg;
    lgam_digam_diff(te, y, lg, dg);
    const float t1 = lgamma_1p(y) - lg;
    const float t2 = (th + y) * L1 + y * (logf(te) - logf(mu + kEps));
    float nb = t1 + t2;
    if (nb != nb) nb = INFINITY;                           // _nan2inf  loss.py:105
    o.gm = th * (mu - y) * rden;
    dth = f + y * rden - dg;
    if (HAS_PI) {
      const float q = 1.0f - pi + kEps;
      nb -= logf(q);                                       // loss.py:130
      dpi = 1.0f / q;
    }
    o.loss = nb;
  }
  o.gm = m_pass ? o.gm : 0.f;
  if (COND_DISP) {
    const bool d_pass = (d_in > 1e-4f) && (d_in < 1e4f);   // DispAct clip mask (network.py:39)
    o.gd = d_pass ? dth * (-expm1f(-d_in)) : 0.f;          // sigmoid(zd) = 1 - exp(-softplus(zd))
  } else {
    o.gd = dth;
  }
  if (HAS_PI) {
    const float s = pi * (1.0f - pi);
    o.loss = fmaf(ridge * pi, pi, o.loss);                 // loss.py:139-140
    o.gp = (dpi + 2.0f * ridge * pi) * s;
  } else {
    o.gp = 0.f;
  }
  return o;
}

// forward-only value
template <bool HAS_PI>
DCA_HD float zinb_elem_loss(float y, float m, float sf, float th, float pi, float ridge) {
  const float mu = m * sf;
  th = fminf(th, 1e6f);
  const float te = th + kEps;
  const float L1 = log1pf(mu / te);
  float l;
  if (HAS_PI && y < 1e-8f) {
    const float z = expf(-th * L1);
    l = -logf(pi + (1.0f - pi) * z + kEps);
  } else {
    float lg, dg;
    lgam_digam_diff(te, y, lg, dg);
    l = lgamma_1p(y) - lg + (th + y) * L1 + y * (logf(te) - logf(mu + kEps));
    if (l != l) l = INFINITY;
    if (HAS_PI) l -= logf(1.0f - pi + kEps);
  }
  if (HAS_PI) l = fmaf(ridge * pi, pi, l);
  return l;
}

}  // namespace zmath
}  // namespace dca
