"""Per-element parity check shared by tests/, __graft_entry__.smoke() and bench.py (pure numpy, no oracle).

The contract (BASELINE.json north_star): bit-exact for nearest-neighbour / integer fields, masks and
indices; `<= 1e-5` RELATIVE for fp32 bilinear and conservative fields.  The relative bound is applied to
every element, not to the field maximum:

    |got - want| <= RTOL * |want| + ATOL_FRAC * max|want|

The absolute floor (1e-7 of the field's scale = one fp32 rounding of a value of that scale) only matters
for elements that are tiny against the field because their terms cancel (rotated winds, random test
fields); on the path's sums of non-negative terms (moisture, snow: many near-zero values) it is far below
RTOL * |want| for anything that is not exactly 0, and exact zeros must come out as exact zeros.
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-5
ATOL_FRAC = 1e-7


def field_errors(got, want, rtol: float = RTOL, atol_frac: float = ATOL_FRAC) -> dict:
    """Element-wise comparison of two arrays of the same size.  Returns
    ok        every element inside rtol * |want| + atol_frac * max|want|
    exact     arrays identical bit for bit
    max_abs   largest |got - want|
    max_rel   largest |got - want| / |want| over the elements with |want| > 1e-3 * max|want|
    max_rel_all  the same over every nonzero element of want
    worst     largest |got - want| / (rtol |want| + atol)   (<= 1 when ok)
    zeros_kept   exact zeros of want that are exact zeros of got (both directions)"""
    g = np.asarray(got).reshape(-1)
    w = np.asarray(want).reshape(-1)
    if g.shape != w.shape:
        raise ValueError(f"shape mismatch {g.shape} vs {w.shape}")
    if g.size == 0:
        return dict(ok=True, exact=True, max_abs=0.0, max_rel=0.0, max_rel_all=0.0, worst=0.0, zeros_kept=True, n=0)
    exact = bool(np.array_equal(g, w))
    g64 = g.astype(np.float64)
    w64 = w.astype(np.float64)
    aw = np.abs(w64)
    scale = float(aw.max())
    d = np.abs(g64 - w64)
    bound = rtol * aw + atol_frac * scale
    finite = bool(np.isfinite(g64).all())
    big = aw > 1e-3 * scale
    nz = aw > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        max_rel = float((d[big] / aw[big]).max()) if big.any() else 0.0
        max_rel_all = float((d[nz] / aw[nz]).max()) if nz.any() else 0.0
        worst = float((d / np.where(bound > 0, bound, 1.0))[bound > 0].max()) if (bound > 0).any() else 0.0
    zeros_kept = bool(np.array_equal(w64 == 0, g64 == 0)) if scale > 0 else bool((g64 == 0).all())
    ok = finite and bool((d <= bound).all())
    return dict(ok=ok, exact=exact, max_abs=float(d.max()), max_rel=max_rel, max_rel_all=max_rel_all, worst=worst,
                zeros_kept=zeros_kept, n=int(g.size))


def assert_field_close(got, want, name: str = "", rtol: float = RTOL, atol_frac: float = ATOL_FRAC) -> dict:
    e = field_errors(got, want, rtol, atol_frac)
    assert e["ok"], (name, {k: e[k] for k in ("max_abs", "max_rel", "worst")})
    return e


def assert_field_exact(got, want, name: str = "") -> None:
    g = np.asarray(got).reshape(-1)
    w = np.asarray(want).reshape(-1)
    assert g.shape == w.shape and np.array_equal(g, w), (name, int((g != w).sum()) if g.shape == w.shape else "shape")


def summarize(results: dict) -> dict:
    """results: name -> field_errors() dict.  The `parity` object of the bench line."""
    fields = len(results)
    return {
        "fields": fields,
        "ok_fields": sum(1 for r in results.values() if r["ok"]),
        "exact_fields": sum(1 for r in results.values() if r["exact"]),
        "max_rel": max((r["max_rel"] for r in results.values()), default=0.0),
        "max_rel_any_nonzero_element": max((r["max_rel_all"] for r in results.values()), default=0.0),
        "worst_over_bound": max((r["worst"] for r in results.values()), default=0.0),
        "tolerance": f"|got-want| <= {RTOL:g}*|want| + {ATOL_FRAC:g}*max|want| per element; nearest/integer fields bit-exact",
        "failed": sorted(k for k, r in results.items() if not r["ok"]),
    }
