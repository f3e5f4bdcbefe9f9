// TEST INFRASTRUCTURE ONLY — compiles flowconductor_b200/csrc/fc_math.cuh (the per-element arithmetic
// the CUDA kernels execute) with g++ so that tests can check those exact statements against the oracle
// on a machine without a GPU.  Never loaded by the product package.
#include "../../flowconductor_b200/csrc/fc_math.cuh"

using namespace fc;

template <int KC>
static void rqs_apply_t(const RqsParams& c, const float* x, const float* params, float* y, float* lad, long n,
                        unsigned* status) {
  for (long i = 0; i < n; ++i) rqs_eval<KC>(c, x[i], params + i * c.P, y[i], lad[i], *status);
}

template <int KC>
static void rqs_backward_t(const RqsParams& c, const float* x, const float* params, const float* gy, const float* gl,
                           float* gx, float* gp, long n) {
  for (long i = 0; i < n; ++i) rqs_backward_elem<KC>(c, x[i], params + i * c.P, gy[i], gl[i], gx[i], gp + i * c.P);
}

extern "C" {

// element-wise: x[n], params[n, P] -> y[n], lad[n] (per element, not reduced)
int hm_rqs_apply(const float* x, const float* params, float* y, float* lad, long n, const fc_rqs_config* cfg,
                 int use_generic, unsigned* status) {
  RqsParams c;
  int rc = make_rqs_params(cfg, c);
  if (rc) return rc;
  *status = 0;
  if (!use_generic && c.K == 8) rqs_apply_t<8>(c, x, params, y, lad, n, status);
  else if (!use_generic && c.K == 16) rqs_apply_t<16>(c, x, params, y, lad, n, status);
  else if (!use_generic && c.K == 5) rqs_apply_t<5>(c, x, params, y, lad, n, status);
  else rqs_apply_t<0>(c, x, params, y, lad, n, status);
  return 0;
}

int hm_rqs_backward(const float* x, const float* params, const float* gy, const float* gl, float* gx, float* gp,
                    long n, const fc_rqs_config* cfg, int use_generic) {
  RqsParams c;
  int rc = make_rqs_params(cfg, c);
  if (rc) return rc;
  if (!use_generic && c.K == 8) rqs_backward_t<8>(c, x, params, gy, gl, gx, gp, n);
  else if (!use_generic && c.K == 16) rqs_backward_t<16>(c, x, params, gy, gl, gx, gp, n);
  else if (!use_generic && c.K == 5) rqs_backward_t<5>(c, x, params, gy, gl, gx, gp, n);
  else rqs_backward_t<0>(c, x, params, gy, gl, gx, gp, n);
  return 0;
}

void hm_affine_apply(const float* x, const float* raw, const float* shift, float* y, float* lad, long n, int activation,
                     int inverse) {
  for (long i = 0; i < n; ++i) affine_eval(x[i], raw[i], shift[i], activation, inverse, y[i], lad[i]);
}

void hm_affine_backward(const float* x, const float* raw, const float* shift, const float* gy, const float* gl,
                        float* gx, float* graw, float* gshift, long n, int activation, int inverse) {
  for (long i = 0; i < n; ++i)
    affine_backward_elem(x[i], raw[i], shift[i], activation, inverse, gy[i], gl[i], gx[i], graw[i], gshift[i]);
}

void hm_sos_apply(const float* x, const float* params, float* y, float* logj, long n, int n_sigmoids) {
  for (long i = 0; i < n; ++i) sos_eval(x[i], params + i * (3 * n_sigmoids + 1), n_sigmoids, y[i], logj[i]);
}

void hm_sos_backward(const float* x, const float* params, const float* gy, const float* gl, float* gx, float* gp,
                     long n, int n_sigmoids) {
  const int P = 3 * n_sigmoids + 1;
  for (long i = 0; i < n; ++i) sos_backward_elem(x[i], params + i * P, n_sigmoids, gy[i], gl[i], gx[i], gp + i * P);
}

void hm_sos_invert(const float* z, const float* params, float* x, float* logj, long n, int n_sigmoids, int iters,
                   float lim) {
  for (long i = 0; i < n; ++i)
    sos_invert(z[i], params + i * (3 * n_sigmoids + 1), n_sigmoids, iters, lim, x[i], logj[i]);
}

// piecewise-linear spline: one element per call, K raw parameters each (runtime-K and unrolled instantiations)
void hm_linspline_apply(const float* x, const float* params, float* y, float* lad, unsigned* status, long n, int k,
                        int tails, float lo, float hi, int inverse, int unrolled) {
  LinSplineParams c;
  c.K = k; c.tails = tails; c.inverse = inverse; c.left = lo; c.right = hi; c.bottom = lo; c.top = hi;
  c.log_k = (float)log((double)k);
  c.inv_w = 1.f / (hi - lo);
  c.inv_h = 1.f / (hi - lo);
  for (long i = 0; i < n; ++i) {
    unsigned st = 0;
    if (unrolled && k == 8) linspline_eval<8>(c, x[i], params + i * k, y[i], lad[i], st);
    else if (unrolled && k == 10) linspline_eval<10>(c, x[i], params + i * k, y[i], lad[i], st);
    else linspline_eval<0>(c, x[i], params + i * k, y[i], lad[i], st);
    status[0] |= st;
  }
}

void hm_linspline_backward(const float* x, const float* params, const float* gy, const float* gl, float* gx, float* gp,
                           long n, int k, int tails, float lo, float hi, int inverse, int unrolled) {
  LinSplineParams c;
  c.K = k; c.tails = tails; c.inverse = inverse; c.left = lo; c.right = hi; c.bottom = lo; c.top = hi;
  c.log_k = (float)log((double)k);
  c.inv_w = 1.f / (hi - lo);
  c.inv_h = 1.f / (hi - lo);
  for (long i = 0; i < n; ++i) {
    if (unrolled && k == 8) linspline_backward_elem<8>(c, x[i], params + i * k, gy[i], gl[i], gx[i], gp + i * k);
    else if (unrolled && k == 10) linspline_backward_elem<10>(c, x[i], params + i * k, gy[i], gl[i], gx[i], gp + i * k);
    else linspline_backward_elem<0>(c, x[i], params + i * k, gy[i], gl[i], gx[i], gp + i * k);
  }
}

// piecewise-quadratic spline: one element per call, params [uw(K) ; uh(K+1 or K-1)]
static QuadSplineParams hm_quad_params(int k, int tails, float lo, float hi, int inverse) {
  QuadSplineParams c;
  c.K = k; c.tails = tails; c.inverse = inverse; c.left = lo; c.right = hi; c.bottom = lo; c.top = hi;
  c.inv_w = 1.f / (hi - lo); c.inv_h = 1.f / (hi - lo); c.min_w = 1e-3f; c.min_h = 1e-3f; c.wh_scale = 1.f;
  return c;
}

void hm_quadspline_apply(const float* x, const float* params, float* y, float* lad, unsigned* status, long n, int k,
                         int tails, float lo, float hi, int inverse, int unrolled) {
  const QuadSplineParams c = hm_quad_params(k, tails, lo, hi, inverse);
  const int P = tails ? 2 * k - 1 : 2 * k + 1;
  for (long i = 0; i < n; ++i) {
    unsigned st = 0;
    if (unrolled && k == 8) quadspline_eval<8>(c, x[i], params + i * P, y[i], lad[i], st);
    else if (unrolled && k == 10) quadspline_eval<10>(c, x[i], params + i * P, y[i], lad[i], st);
    else quadspline_eval<0>(c, x[i], params + i * P, y[i], lad[i], st);
    status[0] |= st;
  }
}

void hm_quadspline_backward(const float* x, const float* params, const float* gy, const float* gl, float* gx, float* gp,
                            long n, int k, int tails, float lo, float hi, int inverse, int unrolled) {
  const QuadSplineParams c = hm_quad_params(k, tails, lo, hi, inverse);
  const int P = tails ? 2 * k - 1 : 2 * k + 1;
  for (long i = 0; i < n; ++i) {
    if (unrolled && k == 8) quadspline_backward_elem<8>(c, x[i], params + i * P, gy[i], gl[i], gx[i], gp + i * P);
    else if (unrolled && k == 10) quadspline_backward_elem<10>(c, x[i], params + i * P, gy[i], gl[i], gx[i], gp + i * P);
    else quadspline_backward_elem<0>(c, x[i], params + i * P, gy[i], gl[i], gx[i], gp + i * P);
  }
}

// cubic spline (element math only so far): params [uw(K) ; uh(K) ; dl ; dr], unit box, no tails
static CubicSplineParams hm_cubic_params(int k, int inverse) {
  CubicSplineParams c;
  c.K = k; c.tails = 0; c.inverse = inverse; c.left = 0.f; c.right = 1.f; c.bottom = 0.f; c.top = 1.f;
  c.inv_w = 1.f; c.inv_h = 1.f; c.min_w = 1e-3f; c.min_h = 1e-3f; c.wh_scale = 1.f;
  return c;
}

void hm_cubicspline_apply(const float* x, const float* params, float* y, float* lad, long n, int k, int inverse,
                          int unrolled) {
  const CubicSplineParams c = hm_cubic_params(k, inverse);
  const int P = 2 * k + 2;
  for (long i = 0; i < n; ++i) {
    unsigned st = 0;
    if (unrolled && k == 8) cubicspline_eval<8>(c, x[i], params + i * P, y[i], lad[i], st);
    else cubicspline_eval<0>(c, x[i], params + i * P, y[i], lad[i], st);
  }
}

void hm_cubicspline_backward(const float* x, const float* params, const float* gy, const float* gl, float* gx, float* gp,
                             long n, int k, int inverse, int unrolled) {
  const CubicSplineParams c = hm_cubic_params(k, inverse);
  const int P = 2 * k + 2;
  for (long i = 0; i < n; ++i) {
    if (unrolled && k == 8) cubicspline_backward_elem<8>(c, x[i], params + i * P, gy[i], gl[i], gx[i], gp + i * P);
    else cubicspline_backward_elem<0>(c, x[i], params + i * P, gy[i], gl[i], gx[i], gp + i * P);
  }
}

// the same with a pre-scale on the raw widths / heights (coupling.py:477-479: / sqrt(hidden))
void hm_cubicspline_scaled(const float* x, const float* params, const float* gy, const float* gl, float* y, float* lad,
                           float* gx, float* gp, long n, int k, int inverse, float scale) {
  CubicSplineParams c = hm_cubic_params(k, inverse);
  c.wh_scale = scale;
  const int P = 2 * k + 2;
  for (long i = 0; i < n; ++i) {
    unsigned st = 0;
    cubicspline_eval<0>(c, x[i], params + i * P, y[i], lad[i], st);
    cubicspline_backward_elem<0>(c, x[i], params + i * P, gy[i], gl[i], gx[i], gp + i * P);
  }
}

// compile-time n = 10 instantiations (what the kernels run for the default sigmoid count)
void hm_sos_apply_n10(const float* x, const float* params, float* y, float* logj, long n) {
  for (long i = 0; i < n; ++i) sos_eval_t<10>(x[i], params + i * 31, 10, y[i], logj[i]);
}

void hm_sos_backward_n10(const float* x, const float* params, const float* gy, const float* gl, float* gx, float* gp,
                         long n) {
  for (long i = 0; i < n; ++i) sos_backward_elem_t<10>(x[i], params + i * 31, 10, gy[i], gl[i], gx[i], gp + i * 31);
}

// the instantiation the kernels use for n = 10 (constants precomputed into registers); also returns the largest
// number of function evaluations any element needed
long hm_sos_invert_n10(const float* z, const float* params, float* x, float* logj, long n, int iters, float lim) {
  long worst = 0;
  for (long i = 0; i < n; ++i) {
    const float* raw = params + i * 31;
    sos_invert_t<10>(z[i], raw, 10, iters, lim, x[i], logj[i]);
    SosConsts<10> k;
    sos_precompute<10>(raw, k);
    long evals = 0;
    sos_solve(z[i], iters, lim, [&](float xx, float& y, float& J) {
      ++evals;
      sos_eval_pre<10>(xx, k, y, J);
    });
    worst = evals > worst ? evals : worst;
  }
  return worst;
}
}
