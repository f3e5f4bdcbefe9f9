"""TEST INFRASTRUCTURE ONLY — CPU restatement of FlowConductor's element-wise bijection hot path.

This file is the parity ORACLE for the sm_100a kernels in `flowconductor_b200/csrc`.  It is never
imported by the product package; only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline /
`--impl reference` legs of `bench.py` may use it.

It restates, in plain CPU PyTorch (dtype-generic: fp32 to mirror the reference, fp64 as ground
truth, autograd-capable for gradient checks), the algorithms of the reference files listed in
SURVEY.md §8(a).  Every function cites the reference file:line it follows (paths relative to
/root/reference).  The operation ORDER of the reference is kept on purpose (theta from knot
differences, log(num) - 2 log(den), root = 2c / (-b - sqrt(disc)), softmax -> cumsum -> rescale ->
forced end knots -> re-derived widths), because fp32 parity depends on it.

Parity is PINNED: `tests/test_oracle_golden.py` checks this file against golden vectors produced
by the unmodified reference (`oracle/make_golden.py`, run in the build container where
/root/reference is importable) and, when the reference tree is present, against the live reference.
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_MIN_BIN_WIDTH = 1e-3  # flowcon/transforms/splines/rational_quadratic.py:8
DEFAULT_MIN_BIN_HEIGHT = 1e-3  # :9
DEFAULT_MIN_DERIVATIVE = 1e-3  # :10


class InputOutsideDomain(Exception):
    """flowcon/transforms/base.py:16-19."""


# --------------------------------------------------------------------------------------------
# a3 / a4 helpers
# --------------------------------------------------------------------------------------------
def bin_index(knots, x, eps=1e-6):
    """flowcon/utils/torchutils.py:147-149 (`searchsorted`): bump the LAST knot by eps, then count
    knots <= x and subtract one.  The reference bumps in place; the bumped knot is never gathered,
    so a copy is equivalent."""
    bumped = knots.clone()
    bumped[..., -1] += eps
    return torch.sum(x[..., None] >= bumped, dim=-1) - 1


def sum_except_batch(x, num_batch_dims=1):
    """flowcon/utils/torchutils.py:25-30."""
    dims = list(range(num_batch_dims, x.dim()))
    return torch.sum(x, dim=dims) if dims else x


# --------------------------------------------------------------------------------------------
# a2: rational_quadratic_spline  (flowcon/transforms/splines/rational_quadratic.py:66-181)
# --------------------------------------------------------------------------------------------
def _knots(unnormalized, lo, hi, min_size):
    """:91-98 (widths) and :106-113 (heights): softmax, floor, cumsum, pad, rescale, force the two
    end knots, re-derive the bin sizes from knot differences."""
    num_bins = unnormalized.shape[-1]
    sizes = F.softmax(unnormalized, dim=-1)
    sizes = min_size + (1 - min_size * num_bins) * sizes
    cum = torch.cumsum(sizes, dim=-1)
    cum = F.pad(cum, pad=(1, 0), mode="constant", value=0.0)
    cum = (hi - lo) * cum + lo
    cum[..., 0] = lo
    cum[..., -1] = hi
    sizes = cum[..., 1:] - cum[..., :-1]
    return cum, sizes


def rational_quadratic_spline(
    inputs,
    unnormalized_widths,
    unnormalized_heights,
    unnormalized_derivatives,
    inverse=False,
    left=0.0,
    right=1.0,
    bottom=0.0,
    top=1.0,
    min_bin_width=DEFAULT_MIN_BIN_WIDTH,
    min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
    min_derivative=DEFAULT_MIN_DERIVATIVE,
    enable_identity_init=False,
    check_domain=True,
):
    """Monotone rational-quadratic spline on [left,right] -> [bottom,top]; returns (outputs,
    per-element logabsdet).  rational_quadratic.py:66-181."""
    if check_domain and inputs.numel() > 0:
        if torch.min(inputs) < left or torch.max(inputs) > right:  # :81-82
            raise InputOutsideDomain()
    num_bins = unnormalized_widths.shape[-1]
    if min_bin_width * num_bins > 1.0:  # :86-89
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * num_bins > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")

    cumwidths, widths = _knots(unnormalized_widths, left, right, min_bin_width)
    beta = math.log(2) / (1 - min_derivative) if enable_identity_init else 1  # :100-103
    derivatives = min_derivative + F.softplus(unnormalized_derivatives, beta=beta)  # :104
    cumheights, heights = _knots(unnormalized_heights, bottom, top, min_bin_height)

    k = bin_index(cumheights if inverse else cumwidths, inputs)[..., None]  # :115-118

    def pick(t):
        return t.gather(-1, k)[..., 0]

    cw, w = pick(cumwidths), pick(widths)  # :120-121
    ch = pick(cumheights)  # :123
    delta = pick(heights / widths)  # :124-125
    d0 = pick(derivatives)  # :127
    d1 = pick(derivatives[..., 1:])  # :128
    h = pick(heights)  # :130

    if inverse:  # :132-160
        u = inputs - ch
        s = d0 + d1 - 2 * delta
        a = u * s + h * (delta - d0)
        b = h * d0 - u * s
        c = -delta * u
        discriminant = b.pow(2) - 4 * a * c
        assert (discriminant >= 0).all()  # :142
        root = (2 * c) / (-b - torch.sqrt(discriminant))
        outputs = root * w + cw
        t1mt = root * (1 - root)
        denominator = delta + s * t1mt
        numerator = delta.pow(2) * (d1 * root.pow(2) + 2 * delta * t1mt + d0 * (1 - root).pow(2))
        logabsdet = torch.log(numerator) - 2 * torch.log(denominator)
        return outputs, -logabsdet
    theta = (inputs - cw) / w  # :162
    t1mt = theta * (1 - theta)
    numerator = h * (delta * theta.pow(2) + d0 * t1mt)
    denominator = delta + (d0 + d1 - 2 * delta) * t1mt
    outputs = ch + numerator / denominator
    dnum = delta.pow(2) * (d1 * theta.pow(2) + 2 * delta * t1mt + d0 * (1 - theta).pow(2))
    logabsdet = torch.log(dnum) - 2 * torch.log(denominator)
    return outputs, logabsdet


# --------------------------------------------------------------------------------------------
# a1: unconstrained_rational_quadratic_spline  (rational_quadratic.py:13-63)
# --------------------------------------------------------------------------------------------
def unconstrained_rational_quadratic_spline(
    inputs,
    unnormalized_widths,
    unnormalized_heights,
    unnormalized_derivatives,
    inverse=False,
    tails="linear",
    tail_bound=1.0,
    min_bin_width=DEFAULT_MIN_BIN_WIDTH,
    min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
    min_derivative=DEFAULT_MIN_DERIVATIVE,
    enable_identity_init=False,
):
    """Linear-tail wrapper.  The reference gathers the inside elements with a boolean mask, calls
    the spline on them and scatters back (:43-61); per element that is the same as evaluating the
    spline everywhere on a clamped copy and selecting, which is what is done here (no data-dependent
    shapes, identical arithmetic on every inside element)."""
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))  # :41
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)  # :26, both ends inclusive
    derivs = F.pad(unnormalized_derivatives, pad=(1, 1))  # :33
    constant = math.log(math.exp(1 - min_derivative) - 1)  # :34
    derivs[..., 0] = constant
    derivs[..., -1] = constant
    safe = torch.where(inside, inputs, torch.zeros_like(inputs))
    y_in, lad_in = rational_quadratic_spline(
        safe,
        unnormalized_widths,
        unnormalized_heights,
        derivs,
        inverse=inverse,
        left=-tail_bound,
        right=tail_bound,
        bottom=-tail_bound,
        top=tail_bound,
        min_bin_width=min_bin_width,
        min_bin_height=min_bin_height,
        min_derivative=min_derivative,
        enable_identity_init=enable_identity_init,
        check_domain=False,
    )
    outputs = torch.where(inside, y_in, inputs)  # :38
    logabsdet = torch.where(inside, lad_in, torch.zeros_like(lad_in))  # :39
    return outputs, logabsdet


def split_rq_params(params, num_bins, wh_divisor=None):
    """flowcon/transforms/coupling.py:549-556 / autoregressive.py:585-591: per feature
    [w_0..w_{K-1} ; h_0..h_{K-1} ; d...]; widths and heights divided by sqrt(hidden) iff the
    conditioner exposes `.hidden_features` (ResidualNet yes, transforms.made.MADE no)."""
    uw = params[..., :num_bins]
    uh = params[..., num_bins : 2 * num_bins]
    ud = params[..., 2 * num_bins :]
    if wh_divisor is not None:
        uw = uw / wh_divisor
        uh = uh / wh_divisor
    return uw, uh, ud


def rq_elementwise(inputs, params, num_bins, tails, tail_bound, inverse, wh_divisor, identity_init,
                   min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
                   min_derivative=DEFAULT_MIN_DERIVATIVE, constrained_bound=None):
    """Shared body of coupling.py:549-582 (`_piecewise_cdf`), autoregressive.py:578-615 and
    conditional.py:700-741: view params [B, D_t, P], slice, optional scaling, dispatch on tails.
    `constrained_bound`: None -> [0,1] (coupling, tails=None); 1.2 -> [-1.2,1.2] (AR/conditional)."""
    b, d = inputs.shape
    params = params.reshape(b, d, -1)
    uw, uh, ud = split_rq_params(params, num_bins, wh_divisor)
    kw = dict(inverse=inverse, min_bin_width=min_bin_width, min_bin_height=min_bin_height,
              min_derivative=min_derivative, enable_identity_init=identity_init)
    if tails is None:
        if constrained_bound is not None:
            c = constrained_bound
            kw.update(left=-c, right=c, bottom=-c, top=c)
        y, lad = rational_quadratic_spline(inputs, uw, uh, ud, **kw)
    else:
        y, lad = unconstrained_rational_quadratic_spline(inputs, uw, uh, ud, tails=tails,
                                                         tail_bound=tail_bound, **kw)
    return y, sum_except_batch(lad)  # coupling.py:293


# --------------------------------------------------------------------------------------------
# n3: piecewise-linear spline  (flowcon/transforms/splines/linear.py:9-105)
# --------------------------------------------------------------------------------------------
def linear_spline(inputs, unnormalized_pdf, inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0):
    """linear.py:38-105.  K equal-width bins with softmax probabilities; forward: bin = floor(u K), value = cdf at the
    bin's left edge + position * probability (:84-95); inverse: bin by searchsorted on the cdf, then the bin's affine
    map inverted through slope / offset (:60-78).  Quirk kept: `searchsorted` bumps the forced last cdf knot (1.0, :56)
    by 1e-6 IN PLACE (torchutils.py:147-149) before the slopes are taken, so the last bin's slope sees the bump."""
    if torch.min(inputs) < left or torch.max(inputs) > right:
        raise InputOutsideDomain()  # :45-46
    lo, hi = (bottom, top) if inverse else (left, right)
    u = (inputs - lo) / (hi - lo)
    k = unnormalized_pdf.size(-1)
    pdf = F.softmax(unnormalized_pdf, dim=-1)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf[..., -1] = 1.0
    cdf = F.pad(cdf, pad=(1, 0), mode="constant", value=0.0)
    if inverse:
        cdf[..., -1] += 1e-6
        idx = (torch.sum(u[..., None] >= cdf, dim=-1) - 1).unsqueeze(-1)
        edges = torch.linspace(0, 1, k + 1).view([1] * u.dim() + [-1]).expand(*u.shape, -1).to(cdf.dtype)
        slopes = (cdf[..., 1:] - cdf[..., :-1]) / (edges[..., 1:] - edges[..., :-1])
        offsets = cdf[..., 1:] - slopes * edges[..., 1:]
        slope = slopes.gather(-1, idx)[..., 0]
        out = torch.clamp((u - offsets.gather(-1, idx)[..., 0]) / slope, 0, 1)
        logabsdet = -torch.log(slope)
        out = out * (right - left) + left
    else:
        pos = u * k
        idx = torch.floor(pos).long()
        idx[idx >= k] = k - 1
        frac = pos - idx.to(u.dtype)
        p = pdf.gather(-1, idx[..., None])[..., 0]
        out = cdf.gather(-1, idx[..., None])[..., 0]
        out = torch.clamp(out + frac * p, 0, 1)
        logabsdet = torch.log(p) - math.log(1.0 / k)
        out = out * (top - bottom) + bottom
    return out, logabsdet


def unconstrained_linear_spline(inputs, unnormalized_pdf, inverse=False, tail_bound=1.0, tails="linear"):
    """linear.py:9-35: identity (logabsdet 0) outside [-tail_bound, tail_bound], both ends inclusive."""
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
    outputs = torch.where(inside, torch.zeros_like(inputs), inputs)
    logabsdet = torch.zeros_like(inputs)
    if torch.any(inside):
        o, l = linear_spline(inputs[inside], unnormalized_pdf[inside, :], inverse=inverse, left=-tail_bound,
                             right=tail_bound, bottom=-tail_bound, top=tail_bound)
        outputs = outputs.masked_scatter(inside, o)
        logabsdet = logabsdet.masked_scatter(inside, l)
    return outputs, logabsdet


def linear_elementwise(inputs, params, num_bins, tails, tail_bound, inverse):
    """coupling.py:340-352 (`PiecewiseLinearCouplingTransform._piecewise_cdf`): params [B, D_t*K] viewed [B, D_t, K],
    no scaling; log-det summed per sample (coupling.py:293)."""
    b, d = inputs.shape
    params = params.reshape(b, d, num_bins)
    if tails is None:
        y, lad = linear_spline(inputs, params, inverse=inverse)
    else:
        y, lad = unconstrained_linear_spline(inputs, params, inverse=inverse, tails=tails, tail_bound=tail_bound)
    return y, sum_except_batch(lad)


# --------------------------------------------------------------------------------------------
# n3: piecewise-quadratic spline  (flowcon/transforms/splines/quadratic.py:11-159)
# --------------------------------------------------------------------------------------------
def _quadratic_knots(raw_w, raw_h, floor_w, floor_h):
    """Bin widths and knot heights of the piecewise-linear pdf, quadratic.py:81-108.
    widths: softmax with a floor (:81-82).  heights: softplus + 1e-3 (:84); when only K-1 raw heights are given the two
    boundary knots get the value that turns into exactly 1 after normalisation (:86-101); then divide by the
    trapezoid area of the un-normalised pdf and apply the height floor (:103-108)."""
    k = raw_w.shape[-1]
    w = floor_w + (1 - floor_w * k) * F.softmax(raw_w, dim=-1)
    e = F.softplus(raw_h) + 1e-3
    if e.shape[-1] == k - 1:
        half_first, half_last = 0.5 * w[..., 0], 0.5 * w[..., -1]
        inner = torch.sum(((e[..., :-1] + e[..., 1:]) / 2) * w[..., 1:-1], dim=-1)
        edge = (0.5 * half_first * e[..., 0] + 0.5 * half_last * e[..., -1] + inner) / (1 - 0.5 * half_first
                                                                                       - 0.5 * half_last)
        edge = edge[..., None]
        e = torch.cat([edge, e, edge], dim=-1)
    area = torch.sum(((e[..., :-1] + e[..., 1:]) / 2) * w, dim=-1)[..., None]
    h = floor_h + (1 - floor_h) * (e / area)
    return w, h


def _padded_cumsum_to_one(x):
    """cumsum with the last entry forced to 1 and a leading 0 (quadratic.py:110-118, linear.py:55-57)."""
    c = torch.cumsum(x, dim=-1)
    c[..., -1] = 1.0
    return F.pad(c, pad=(1, 0), mode="constant", value=0.0)


def quadratic_spline(inputs, unnormalized_widths, unnormalized_heights, inverse=False, left=0.0, right=1.0, bottom=0.0,
                     top=1.0, min_bin_width=1e-3, min_bin_height=1e-3):
    """quadratic.py:55-159: piecewise-linear pdf (K widths, K+1 knot heights) integrated to a piecewise-quadratic cdf.
    Operation order of the reference kept: normalise the input (:70-73), knots (:81-108), left-cdf and location
    vectors (:110-118), bin by searchsorted on the cdf (inverse) or the locations (forward) (:120-123), per-bin
    polynomial a alpha^2 + b alpha + c (:130-132), root / evaluation (:134-149), rescale (:151-154)."""
    if torch.min(inputs) < left or torch.max(inputs) > right:
        raise InputOutsideDomain()  # :67-68
    lo, hi = (bottom, top) if inverse else (left, right)
    u = (inputs - lo) / (hi - lo)
    k = unnormalized_widths.shape[-1]
    if min_bin_width * k > 1.0:
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * k > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")
    w, h = _quadratic_knots(unnormalized_widths, unnormalized_heights, min_bin_width, min_bin_height)
    left_cdf = _padded_cumsum_to_one(((h[..., :-1] + h[..., 1:]) / 2) * w)
    locations = _padded_cumsum_to_one(w)
    idx = bin_index(left_cdf if inverse else locations, u)[..., None]
    loc = locations.gather(-1, idx)[..., 0]
    bw = w.gather(-1, idx)[..., 0]
    c = left_cdf.gather(-1, idx)[..., 0]
    h_left = h.gather(-1, idx)[..., 0]
    h_right = h.gather(-1, idx + 1)[..., 0]
    a = 0.5 * (h_right - h_left) * bw
    b = h_left * bw
    if inverse:
        alpha = (-b + torch.sqrt(b.pow(2) - 4 * a * (c - u))) / (2 * a)
        out = torch.clamp(alpha * bw + loc, 0, 1)
        logabsdet = -torch.log(alpha * (h_right - h_left) + h_left)
        out = out * (right - left) + left
    else:
        alpha = (u - loc) / bw
        out = torch.clamp(a * alpha.pow(2) + b * alpha + c, 0, 1)
        logabsdet = torch.log(alpha * (h_right - h_left) + h_left)
        out = out * (top - bottom) + bottom
    return out, logabsdet


def unconstrained_quadratic_spline(inputs, unnormalized_widths, unnormalized_heights, inverse=False, tail_bound=1.0,
                                   tails="linear", min_bin_width=1e-3, min_bin_height=1e-3):
    """quadratic.py:11-52: identity outside [-tail_bound, tail_bound]; K-1 raw heights."""
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    assert unnormalized_heights.shape[-1] == unnormalized_widths.shape[-1] - 1
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
    outputs = torch.where(inside, torch.zeros_like(inputs), inputs)
    logabsdet = torch.zeros_like(inputs)
    if torch.any(inside):
        o, l = quadratic_spline(inputs[inside], unnormalized_widths[inside, :], unnormalized_heights[inside, :],
                                inverse=inverse, left=-tail_bound, right=tail_bound, bottom=-tail_bound, top=tail_bound,
                                min_bin_width=min_bin_width, min_bin_height=min_bin_height)
        outputs = outputs.masked_scatter(inside, o)
        logabsdet = logabsdet.masked_scatter(inside, l)
    return outputs, logabsdet


# --------------------------------------------------------------------------------------------
# n3 (groundwork for the next round: no kernel yet): cubic spline  (flowcon/transforms/splines/cubic.py:15-267)
# --------------------------------------------------------------------------------------------
def _cbrt(x):
    """flowcon/utils/torchutils.py:152-154."""
    return torch.sign(x) * torch.exp(torch.log(torch.abs(x)) / 3.0)


def _cubic_coefficients(raw_w, raw_h, raw_dl, raw_dr, floor_w, floor_h):
    """Knots and per-bin cubic coefficients, cubic.py:98-138: floored softmax widths / heights and their forced-to-one
    cumulative sums (:98-110); bin slopes; interior knot derivatives by the monotone (Steffen-type) rule
    min(|s_i|, |s_{i+1}|, weighted mean) * (sign s_i + sign s_{i+1}) (:112-132); boundary derivatives
    sigmoid(raw) * 3 * slope (:121-126); a, b, c, d of a t^3 + b t^2 + c t + d with t measured from the bin's left edge."""
    k = raw_w.shape[-1]
    w = floor_w + (1 - floor_w * k) * F.softmax(raw_w, dim=-1)
    h = floor_h + (1 - floor_h * k) * F.softmax(raw_h, dim=-1)
    cum_w = _padded_cumsum_to_one(w)
    cum_h = _padded_cumsum_to_one(h)
    slope = h / w
    bound_abs = torch.min(torch.abs(slope[..., :-1]), torch.abs(slope[..., 1:]))
    bound_mean = 0.5 * (w[..., 1:] * slope[..., :-1] + w[..., :-1] * slope[..., 1:]) / (w[..., :-1] + w[..., 1:])
    inner = torch.min(bound_abs, bound_mean) * (torch.sign(slope[..., :-1]) + torch.sign(slope[..., 1:]))
    d_left = torch.sigmoid(raw_dl) * 3 * slope[..., 0][..., None]
    d_right = torch.sigmoid(raw_dr) * 3 * slope[..., -1][..., None]
    deriv = torch.cat([d_left, inner, d_right], dim=-1)
    a = (deriv[..., :-1] + deriv[..., 1:] - 2 * slope) / w.pow(2)
    b = (3 * slope - 2 * deriv[..., :-1] - deriv[..., 1:]) / w
    return cum_w, cum_h, a, b, deriv[..., :-1], cum_h[..., :-1]


def _cubic_root_in_bin(a, b, c, d, target, lo, hi, eps, quadratic_threshold):
    """Root of a t^3 + b t^2 + c t + d = target inside [lo, hi] (absolute position = t + lo), cubic.py:152-237 (Blinn
    2007): one real root -> Cardano with cube roots (:176-187); three real roots -> trigonometric form, pick the root
    inside the bin (first one whose [lo - eps, hi + eps] test passes, via argsort of the masks, :191-225); |a| below the
    threshold -> the quadratic's root (:229-234, applied last, overriding)."""
    b3 = (b / a) / 3.0
    c3 = (c / a) / 3.0
    d0 = (d - target) / a
    delta_1 = -b3.pow(2) + c3
    delta_2 = -c3 * b3 + d0
    delta_3 = b3 * d0 - c3.pow(2)
    disc = 4.0 * delta_1 * delta_3 - delta_2.pow(2)
    dep_1 = -2.0 * b3 * delta_1 + delta_2
    three = disc >= 0
    one = disc < 0
    out = torch.zeros_like(target)
    sq = torch.sqrt(-disc[one])
    out[one] = _cbrt((-dep_1[one] + sq) / 2.0) + _cbrt((-dep_1[one] - sq) / 2.0) - b3[one] + lo[one]
    theta = torch.atan2(torch.sqrt(disc[three]), -dep_1[three]) / 3.0
    cos_t, sin_t = torch.cos(theta), torch.sin(theta)
    scale = 2 * torch.sqrt(-delta_1[three])
    shift = -b3[three] + lo[three]
    cands = torch.stack([cos_t, -0.5 * cos_t - 0.5 * math.sqrt(3) * sin_t, -0.5 * cos_t + 0.5 * math.sqrt(3) * sin_t],
                        dim=-1) * scale[..., None] + shift[..., None]
    ok = ((lo[three][..., None] - eps) < cands).to(target.dtype) * (cands < (hi[three][..., None] + eps)).to(target.dtype)
    pick = torch.argsort(ok, dim=-1, descending=True)[..., 0][..., None]
    out[three] = torch.gather(cands, dim=-1, index=pick).view(-1)
    near_quadratic = a.abs() < quadratic_threshold
    qa, qb, qc = b[near_quadratic], c[near_quadratic], d[near_quadratic] - target[near_quadratic]
    out[near_quadratic] = (-qb + torch.sqrt(qb.pow(2) - 4 * qa * qc)) / (2 * qa) + lo[near_quadratic]
    return out


def unconstrained_cubic_spline(inputs, unnormalized_widths, unnormalized_heights, unnorm_derivatives_left,
                               unnorm_derivatives_right, inverse=False, tail_bound=1.0, tails="linear"):
    """cubic.py:15-60: identity (logabsdet 0) outside [-tail_bound, tail_bound]."""
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
    outputs = torch.where(inside, torch.zeros_like(inputs), inputs)
    logabsdet = torch.zeros_like(inputs)
    if torch.any(inside):
        o, l = cubic_spline(inputs[inside], unnormalized_widths[inside, :], unnormalized_heights[inside, :],
                            unnorm_derivatives_left[inside, :], unnorm_derivatives_right[inside, :], inverse=inverse,
                            left=-tail_bound, right=tail_bound, bottom=-tail_bound, top=tail_bound)
        outputs = outputs.masked_scatter(inside, o)
        logabsdet = logabsdet.masked_scatter(inside, l)
    return outputs, logabsdet


def cubic_elementwise(inputs, params, num_bins, tails, tail_bound, inverse, divisor):
    """coupling.py:468-500: params [B, D_t * (2K+2)] viewed [B, D_t, 2K+2]; widths and heights / sqrt(hidden)."""
    b, d = inputs.shape
    params = params.reshape(b, d, -1)
    uw, uh = params[..., :num_bins] / divisor, params[..., num_bins : 2 * num_bins] / divisor
    dl, dr = params[..., 2 * num_bins][..., None], params[..., 2 * num_bins + 1][..., None]
    if tails is None:
        flat = [t.reshape(b * d, -1) for t in (uw, uh, dl, dr)]
        y, lad = cubic_spline(inputs.reshape(b * d), *flat, inverse=inverse)
        y, lad = y.reshape(b, d), lad.reshape(b, d)
    else:
        y, lad = unconstrained_cubic_spline(inputs, uw, uh, dl, dr, inverse=inverse, tails=tails, tail_bound=tail_bound)
    return y, sum_except_batch(lad)


def cubic_cdf(state, prefix, inputs, num_bins, tails, tail_bound, inverse):
    """PiecewiseCubicCDF._spline, nonlinearities.py:363-398: parameters shared across the batch."""
    b = inputs.shape[0]
    ps = [state[prefix + n].to(inputs.dtype)[None].expand(b, -1, -1) for n in
          ("unnormalized_widths", "unnormalized_heights", "unnorm_derivatives_left", "unnorm_derivatives_right")]
    if tails is None:
        d = inputs.shape[1]
        y, lad = cubic_spline(inputs.reshape(b * d), *[t.reshape(b * d, -1) for t in ps], inverse=inverse)
        y, lad = y.reshape(b, d), lad.reshape(b, d)
    else:
        y, lad = unconstrained_cubic_spline(inputs, *ps, inverse=inverse, tails=tails, tail_bound=tail_bound)
    return y, sum_except_batch(lad)


def cubic_spline(inputs, unnormalized_widths, unnormalized_heights, unnorm_derivatives_left, unnorm_derivatives_right,
                 inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0, min_bin_width=1e-3, min_bin_height=1e-3,
                 eps=1e-5, quadratic_threshold=1e-3):
    """cubic.py:63-267 (inputs are flat [n] with [n, K] parameters, as the reference's masked call sites pass them)."""
    if torch.min(inputs) < left or torch.max(inputs) > right:
        raise InputOutsideDomain()
    k = unnormalized_widths.shape[-1]
    if min_bin_width * k > 1.0:
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * k > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")
    lo, hi = (bottom, top) if inverse else (left, right)
    u = (inputs - lo) / (hi - lo)
    cum_w, cum_h, a, b, c, d = _cubic_coefficients(unnormalized_widths, unnormalized_heights, unnorm_derivatives_left,
                                                   unnorm_derivatives_right, min_bin_width, min_bin_height)
    idx = bin_index(cum_h if inverse else cum_w, u)[..., None]
    ai, bi, ci, di = (t.gather(-1, idx)[..., 0] for t in (a, b, c, d))
    x_lo = cum_w.gather(-1, idx)[..., 0]
    x_hi = cum_w.gather(-1, idx + 1)[..., 0]
    if inverse:
        out = _cubic_root_in_bin(ai, bi, ci, di, u, x_lo, x_hi, eps, quadratic_threshold)
        t = out - x_lo
        logabsdet = -torch.log(3 * ai * t.pow(2) + 2 * bi * t + ci)
        out = out * (right - left) + left
    else:
        t = u - x_lo
        out = ai * t.pow(3) + bi * t.pow(2) + ci * t + di
        logabsdet = torch.log(3 * ai * t.pow(2) + 2 * bi * t + ci)
        out = out * (top - bottom) + bottom
    return out, logabsdet


# --------------------------------------------------------------------------------------------
# a7 / a9: affine element-wise transforms
# --------------------------------------------------------------------------------------------
def affine_scale(unconstrained, activation):
    """'sigmoid2': coupling.py:224 default; 'softplus_clamp3': coupling.py:225 general;
    'softplus_eps': autoregressive.py:102 (MAF, epsilon 1e-3 from :91)."""
    if activation == "sigmoid2":
        return torch.sigmoid(unconstrained + 2) + 1e-3
    if activation == "softplus_clamp3":
        return (F.softplus(unconstrained) + 1e-3).clamp(0, 3)
    if activation == "softplus_eps":
        return F.softplus(unconstrained) + 1e-3
    raise ValueError(activation)


def affine_elementwise(inputs, params, layout, activation, inverse):
    """coupling.py:234-252 (layout 'blocked': params[:, :D_t] = shift, params[:, D_t:] = raw scale)
    and autoregressive.py:97-129 (layout 'interleaved': view [B,D,2], [...,0] raw scale, [...,1]
    shift)."""
    b, d = inputs.shape
    if layout == "blocked":
        shift, raw = params[:, :d], params[:, d:]
    elif layout == "interleaved":
        p = params.reshape(b, d, 2)
        raw, shift = p[..., 0], p[..., 1]
    else:
        raise ValueError(layout)
    scale = affine_scale(raw, activation)
    log_scale = torch.log(scale)
    if inverse:
        return (inputs - shift) / scale, -sum_except_batch(log_scale)
    return inputs * scale + shift, sum_except_batch(log_scale)


# --------------------------------------------------------------------------------------------
# a11 / a12: sum of sigmoids + extended softplus
# --------------------------------------------------------------------------------------------
SOS_SCALE_MIN = 0.1  # flowcon/transforms/adaptive_sigmoids.py:22
SOS_SCALE_MAX = 10.0  # :23
SOS_SHIFT_MAX = 10  # :24
SOS_EPS = 1e-6  # :69


def sos_forward_elementwise(inputs, raw_params, n_sigmoids):
    """adaptive_sigmoids.py:111-142 with raw params [B, D, 3n+1] split as :92, and
    ExtendedSoftplus flowcon/transforms/nonlinearities.py:519-552.  Returns (outputs [B,D],
    per-element log-derivative [B,D])."""
    n = n_sigmoids
    shift_raw, logscale_raw, softmax_raw, esp_raw = torch.split(raw_params, [n, n, n, 1], dim=-1)
    # get_params :132-142  (log_scale_postact == 0 -> exp() == 1, :67)
    weights = F.softmax(softmax_raw, dim=-1) + SOS_EPS
    weights = weights / weights.sum(-1).unsqueeze(-1)
    weights = math.exp(0.0) * weights
    scale = torch.sigmoid(logscale_raw) * (SOS_SCALE_MAX - SOS_SCALE_MIN) + SOS_SCALE_MIN
    shift = torch.tanh(shift_raw) * SOS_SHIFT_MAX
    # sum_of_sigmoids :120-130
    pre = scale * (inputs.unsqueeze(-1) - shift)
    sig = weights * torch.sigmoid(pre)
    log_jac = torch.log(weights) + torch.log(scale) + (pre - 2 * F.softplus(pre))  # :108-109
    y_sig = sig.sum(-1) / weights.sum(-1)
    logj_sig = torch.logsumexp(log_jac, -1)
    # ExtendedSoftplus.forward nonlinearities.py:543-552
    s = F.softplus(esp_raw.reshape(inputs.shape)) + 1e-1  # get_shift :519-520
    y_esp = F.softplus(inputs - s) + (-F.softplus(-(inputs + s)))
    logj_esp = torch.logaddexp(-torch.logaddexp(s, inputs) + inputs, -F.softplus(s + inputs))
    return y_sig + y_esp, torch.logaddexp(logj_sig, logj_esp)  # :114-116


def sos_forward(inputs, params, n_sigmoids, offset=0.0):
    """SumOfSigmoids.forward + wrapper offset (autoregressive.py:309 subtracts 0.5; conditional.py
    :774-780 does not).  Returns (outputs, logabsdet[B])."""
    b, d = inputs.shape
    y, logj = sos_forward_elementwise(inputs, params.reshape(b, d, 3 * n_sigmoids + 1), n_sigmoids)
    return y + offset, logj.sum(-1)


def sos_inverse(z, params, n_sigmoids, offset=0.0, num_iterations=50, lim=120.0, ratio_multiplier=1.5,
                atol=1e-7):
    """MonotonicTransform.inverse -> newton_inverse -> bisection_inverse,
    flowcon/transforms/no_analytic_inv/base.py:23-83,100-103, with SoS settings
    (adaptive_sigmoids.py:26,57: 50 iterations, lim 120).  `offset` as in sos_forward: the wrapper
    adds 0.5 to the inputs before inverting (autoregressive.py:313)."""
    b, d = z.shape
    raw = params.reshape(b, d, 3 * n_sigmoids + 1)
    z = z - offset

    def fwd(x):
        return sos_forward_elementwise(x, raw, n_sigmoids)

    def diffs(z_max, z_min):  # calc_diffs :85-92
        dmax = z - z_max
        imax = torch.argmax(dmax)
        dmin = z - z_min
        imin = torch.argmin(dmin)
        return imax, imin, dmax.flatten()[imax], dmin.flatten()[imin]

    with torch.no_grad():
        x_max = torch.ones_like(z) * lim
        x_min = -torch.ones_like(z) * lim
        z_max, _ = fwd(x_max)
        z_min, _ = fwd(x_min)
        imax, imin, maxdiff, mindiff = diffs(z_max, z_min)
        while maxdiff > 0:  # :48-52
            ratio = (maxdiff + z_max.flatten()[imax]) / z_max.flatten()[imax]
            x_max = x_max * ratio_multiplier * ratio
            z_max, _ = fwd(x_max)
            imax, imin, maxdiff, mindiff = diffs(z_max, z_min)
        x_max = x_max + 1
        while mindiff < 0:  # :55-59
            ratio = (mindiff + z_min.flatten()[imin]) / z_min.flatten()[imin]
            x_min = x_min * ratio_multiplier * ratio
            z_min, _ = fwd(x_min)
            imax, imin, maxdiff, mindiff = diffs(z_max, z_min)
        x_min = x_min - 1
        i = 0
        x_mid = (x_max + x_min) / 2
        while i < num_iterations and (x_mid - z).abs().max() > atol:  # :67 (guard compares x to z: quirk)
            x_mid = (x_max + x_min) / 2
            z_mid, _ = fwd(x_mid)
            go_left = (z_mid > z).to(z.dtype)
            go_right = (z_mid < z).to(z.dtype)
            equal = 1 - (go_left + go_right)
            x_max = go_left * x_mid + go_right * x_max + equal * x_mid
            x_min = go_right * x_mid + go_left * x_min + equal * x_mid
            i += 1
        x = (x_max + x_min) / 2
    # two Newton steps, derivative through autograd, df + 1e-7  (:27-33)
    with torch.enable_grad():
        guess = x.detach().requires_grad_(True)
        for _ in range(2):
            f = fwd(guess)[0] - z
            df = torch.autograd.grad(f, [guess], grad_outputs=torch.ones_like(f), create_graph=True)[0]
            guess = guess - f / (df + 1e-7)
    _, logj = fwd(guess)
    return guess, -logj.sum(-1)


# --------------------------------------------------------------------------------------------
# a15: conditioner networks evaluated from a state_dict (keys as in the reference modules)
# --------------------------------------------------------------------------------------------
def residual_net(state, prefix, inputs, context=None, num_blocks=2):
    """flowcon/nn/nets/resnet.py:55-100 (ResidualNet) with ResidualBlock :9-52, relu activation,
    no dropout / batch norm (the configurations on the path)."""
    def lin(name, t):
        return F.linear(t, state[prefix + name + ".weight"], state[prefix + name + ".bias"])

    temps = lin("initial_layer", inputs if context is None else torch.cat((inputs, context), dim=1))
    for i in range(num_blocks):
        blk = "blocks.{}.".format(i)
        t = F.relu(temps)
        t = lin(blk + "linear_layers.0", t)
        t = F.relu(t)
        t = lin(blk + "linear_layers.1", t)
        if context is not None:
            t = F.glu(torch.cat((t, lin(blk + "context_layer", context)), dim=1), dim=1)
        temps = temps + t
    return lin("final_layer", temps)


def made_net(state, prefix, inputs, context=None, num_blocks=2):
    """flowcon/transforms/made.py:274-283 (MADE.forward, residual blocks :190-202) with
    MaskedLinear.forward :71-72 = F.linear(x, weight * mask, bias)."""
    def mlin(name, t):
        return F.linear(t, state[prefix + name + ".weight"] * state[prefix + name + ".mask"],
                        state[prefix + name + ".bias"])

    def lin(name, t):
        return F.linear(t, state[prefix + name + ".weight"], state[prefix + name + ".bias"])

    temps = mlin("initial_layer", inputs)
    if context is not None:
        temps = temps + F.relu(lin("context_layer", context))
    for i in range(num_blocks):
        blk = "blocks.{}.".format(i)
        t = F.relu(temps)
        t = mlin(blk + "linear_layers.0", t)
        if context is not None:
            t = t + lin(blk + "context_layer", context)
        t = F.relu(t)
        t = mlin(blk + "linear_layers.1", t)
        temps = temps + t
    return mlin("final_layer", temps)


# --------------------------------------------------------------------------------------------
# a5-a10, a14, a16: layers, composition, base density — driven by a plain-data layer spec
# --------------------------------------------------------------------------------------------
def rq_cdf(state, prefix, inputs, num_bins, tails, tail_bound, inverse):
    """PiecewiseRationalQuadraticCDF._spline, flowcon/transforms/nonlinearities.py:451-481: learnable spline
    parameters [D, K] / [D, K] / [D, K-1 or K+1] shared across the batch (`_share_across_batch` :246-247), no 1/sqrt(H)
    scaling, identity-init off, constrained domain [0, 1] when tails is None."""
    b = inputs.shape[0]
    uw = state[prefix + "unnormalized_widths"].to(inputs.dtype)[None].expand(b, -1, -1)
    uh = state[prefix + "unnormalized_heights"].to(inputs.dtype)[None].expand(b, -1, -1)
    ud = state[prefix + "unnormalized_derivatives"].to(inputs.dtype)[None].expand(b, -1, -1)
    if tails is None:
        y, lad = rational_quadratic_spline(inputs, uw, uh, ud, inverse=inverse)
    else:
        y, lad = unconstrained_rational_quadratic_spline(inputs, uw, uh, ud, inverse=inverse, tails=tails,
                                                         tail_bound=tail_bound)
    return y, sum_except_batch(lad)


def quadratic_elementwise(inputs, params, num_bins, tails, tail_bound, inverse):
    """coupling.py:403-427 / autoregressive.py:417-450: params [B, D_t * P] viewed [B, D_t, P], widths first."""
    b, d = inputs.shape
    params = params.reshape(b, d, -1)
    uw, uh = params[..., :num_bins], params[..., num_bins:]
    if tails is None:
        y, lad = quadratic_spline(inputs, uw, uh, inverse=inverse)
    else:
        y, lad = unconstrained_quadratic_spline(inputs, uw, uh, inverse=inverse, tails=tails, tail_bound=tail_bound)
    return y, sum_except_batch(lad)


def quadratic_cdf(state, prefix, inputs, num_bins, tails, tail_bound, inverse):
    """PiecewiseQuadraticCDF._spline, nonlinearities.py:309-334: parameters shared across the batch."""
    b = inputs.shape[0]
    uw = state[prefix + "unnormalized_widths"].to(inputs.dtype)[None].expand(b, -1, -1)
    uh = state[prefix + "unnormalized_heights"].to(inputs.dtype)[None].expand(b, -1, -1)
    if tails is None:
        y, lad = quadratic_spline(inputs, uw, uh, inverse=inverse)
    else:
        y, lad = unconstrained_quadratic_spline(inputs, uw, uh, inverse=inverse, tails=tails, tail_bound=tail_bound)
    return y, sum_except_batch(lad)


def linear_cdf(state, prefix, inputs, num_bins, tails, tail_bound, inverse):
    """PiecewiseLinearCDF._spline, flowcon/transforms/nonlinearities.py:263-277: `unnormalized_pdf` [D, K] shared across
    the batch."""
    u = state[prefix + "unnormalized_pdf"].to(inputs.dtype)[None].expand(inputs.shape[0], -1, -1)
    if tails is None:
        y, lad = linear_spline(inputs, u, inverse=inverse)
    else:
        y, lad = unconstrained_linear_spline(inputs, u, inverse=inverse, tails=tails, tail_bound=tail_bound)
    return y, sum_except_batch(lad)


def _coupling(state, spec, inputs, context, inverse, elementwise):
    """CouplingTransform.forward/inverse, flowcon/transforms/coupling.py:73-130.  With an unconditional transform
    (spec["unconditional"], coupling.py:90-94 / :116-120) the conditioner always sees the identity features on the
    DATA side: forward transforms them after the conditioner ran, inverse before."""
    p = spec["prefix"]
    idf = state[p + "identity_features"]
    trf = state[p + "transform_features"]
    identity = inputs[:, idf]
    transform = inputs[:, trf]
    uncond = spec.get("unconditional", False)
    cdf = {"plin_coupling": linear_cdf, "pquad_coupling": quadratic_cdf, "pcubic_coupling": cubic_cdf}.get(
        spec["kind"], rq_cdf)
    lad_id = 0.0
    if uncond and inverse:
        identity, lad_id = cdf(state, p + "unconditional_transform.", identity, spec["num_bins"], spec.get("tails"),
                               spec.get("tail_bound", 1.0), True)
    params = residual_net(state, p + "transform_net.", identity, context, spec.get("num_blocks", 2))
    transform, lad = elementwise(transform, params)
    if uncond and not inverse:
        identity, lad_id = cdf(state, p + "unconditional_transform.", identity, spec["num_bins"], spec.get("tails"),
                               spec.get("tail_bound", 1.0), False)
    lad = lad + lad_id
    outputs = torch.empty_like(inputs)
    outputs[:, idf] = identity
    outputs[:, trf] = transform
    return outputs, lad


def _autoregressive(state, spec, inputs, context, inverse, elementwise):
    """AutoregressiveTransform.forward/inverse, autoregressive.py:39-53 (inverse = D passes)."""
    p = spec["prefix"] + "autoregressive_net."
    nb = spec.get("num_blocks", 2)
    if not inverse:
        return elementwise(inputs, made_net(state, p, inputs, context, nb))
    outputs = torch.zeros_like(inputs)
    lad = None
    for _ in range(inputs.shape[1]):
        outputs, lad = elementwise(inputs, made_net(state, p, outputs, context, nb))
    return outputs, lad


def apply_layer(state, spec, inputs, context=None, inverse=False):
    """One bijection layer in the direction asked; returns (outputs, logabsdet[B])."""
    kind = spec["kind"]
    p = spec["prefix"]
    if kind == "permutation":  # flowcon/transforms/permutations.py:27-46
        perm = state[p + "_permutation"]
        if inverse:
            perm = torch.argsort(perm)
        return torch.index_select(inputs, 1, perm), inputs.new_zeros(inputs.shape[0])
    if kind == "actnorm":  # flowcon/transforms/normalization.py:173-201 (2-D inputs, eval mode)
        log_scale, shift = state[p + "log_scale"], state[p + "shift"]
        ones = inputs.new_ones(inputs.shape[0])
        if inverse:
            return (inputs - shift.view(1, -1)) / torch.exp(log_scale).view(1, -1), -torch.sum(log_scale) * ones
        return torch.exp(log_scale).view(1, -1) * inputs + shift.view(1, -1), torch.sum(log_scale) * ones
    if kind in ("prq_coupling", "maf_prq", "cond_prq"):
        hidden = spec["hidden_features"]
        # 1/sqrt(H) scaling only where the conditioner exposes .hidden_features (coupling.py:554,
        # conditional.py:711); transforms.made.MADE does not (autoregressive.py:589)
        divisor = None if kind == "maf_prq" else math.sqrt(hidden)
        ident = kind != "prq_coupling"  # autoregressive.py:611, conditional.py:733
        bound = None if kind == "prq_coupling" else 1.2  # autoregressive.py:595, conditional.py:717

        def ew(x, params):
            return rq_elementwise(x, params, spec["num_bins"], spec.get("tails"), spec.get("tail_bound", 1.0),
                                  inverse, divisor, ident, constrained_bound=bound)
    elif kind == "pcubic_coupling":
        def ew(x, params):
            return cubic_elementwise(x, params, spec["num_bins"], spec.get("tails"), spec.get("tail_bound", 1.0), inverse,
                                     math.sqrt(spec["hidden_features"]))
    elif kind == "maf_pcubic":
        # autoregressive.py:492-517: always the constrained unit box; the /sqrt(hidden) of :507-509 never happens
        # (transforms.made.MADE has no `hidden_features` attribute)
        def ew(x, params):
            return cubic_elementwise(x, params, spec["num_bins"], None, 1.0, inverse, 1.0)
    elif kind in ("pquad_coupling", "maf_pquad"):
        # coupling.py:409-411: widths and heights / sqrt(hidden) when the conditioner exposes it (ResidualNet yes, MADE no)
        div = math.sqrt(spec["hidden_features"]) if kind == "pquad_coupling" else 1.0

        def ew(x, params):
            return quadratic_elementwise(x, params / div, spec["num_bins"], spec.get("tails"),
                                         spec.get("tail_bound", 1.0), inverse)
    elif kind in ("plin_coupling", "maf_plin"):
        def ew(x, params):
            return linear_elementwise(x, params, spec["num_bins"], spec.get("tails"), spec.get("tail_bound", 1.0),
                                      inverse)
    elif kind == "affine_coupling":
        def ew(x, params):
            return affine_elementwise(x, params, "blocked", spec.get("scale_activation", "sigmoid2"), inverse)
    elif kind == "maf_affine":
        def ew(x, params):
            return affine_elementwise(x, params, "interleaved", "softplus_eps", inverse)
    elif kind in ("maf_sos", "cond_sos"):
        off = -0.5 if kind == "maf_sos" else 0.0  # autoregressive.py:309,313

        def ew(x, params):
            if inverse:
                return sos_inverse(x, params, spec["n_sigmoids"], offset=off)
            return sos_forward(x, params, spec["n_sigmoids"], offset=off)
    else:
        raise ValueError(kind)

    if kind.endswith("_coupling"):
        return _coupling(state, spec, inputs, context, inverse, ew)
    if kind.startswith("maf_"):
        return _autoregressive(state, spec, inputs, context, inverse, ew)
    # ConditionalTransform.forward/inverse, flowcon/transforms/conditional.py:74-86
    if context is None:
        raise TypeError("Conditional transforms require a context.")
    params = residual_net(state, p + "conditional_net.", context, None, spec.get("num_blocks", 2))
    return ew(inputs, params)


def composite(state, specs, inputs, context=None, inverse=False):
    """CompositeTransform._cascade, flowcon/transforms/base.py:44-60 (inverse walks the layers in
    reverse order)."""
    total = inputs.new_zeros(inputs.shape[0])
    outputs = inputs
    for spec in (reversed(specs) if inverse else specs):
        outputs, lad = apply_layer(state, spec, outputs, context, inverse)
        total = total + lad
    return outputs, total


def standard_normal_log_prob(inputs):
    """StandardNormal._log_prob, flowcon/distributions/normal.py:23-33 (log_z is a 0-dim fp64
    buffer; subtracting it does not promote an fp32 result)."""
    d = inputs.shape[1]
    log_z = torch.tensor(0.5 * d * math.log(2 * math.pi), dtype=torch.float64)
    return -0.5 * sum_except_batch(inputs ** 2) - log_z.to(inputs.dtype)


def flow_log_prob(state, specs, inputs, context=None):
    """Flow._log_prob, flowcon/flows/base.py:41-48."""
    noise, lad = composite(state, specs, inputs, context, inverse=False)
    return standard_normal_log_prob(noise) + lad


def flow_sample_from_noise(state, specs, noise, context=None):
    """Flow._sample, flowcon/flows/base.py:50-74, with the base-distribution draw factored out so
    the same noise can be fed to both implementations."""
    samples, _ = composite(state, specs, noise, context, inverse=True)
    return samples
