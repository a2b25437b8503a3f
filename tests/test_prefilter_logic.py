"""Host-side check of the tracer prefilter's decision rule (csrc/tracer.cu: prefilter_sampler_kernel,
prefilter_argmin_kernel) -- no GPU needed.

The rule is restated here in numpy next to the selection logic of ray_sampler (ray_tracing.py:221-249) and
minimal_sdf_points (:297-300).  Property: for ANY screening values within tau of the exact ones, running the selection on
"exact where the rule asks for a refinement, screening elsewhere" picks the same samples and reads the same values as
running it on the exact values -- including adversarial screening errors of +-0.999 tau, values sitting on +-tau, exact
zeros and python's [-1] wrap at index 0.  The GPU tests then show that the kernels implement this rule bit for bit
(tests/test_gpu_prefilter.py)."""
import numpy as np
import pytest

S = 100


def candidates_sampler(lp, tau, inside_true):
    """prefilter_sampler_kernel: which of the S samples of one ray are re-evaluated exactly."""
    want = np.zeros(S, dtype=bool)
    k = j = -1
    for i in range(S):
        v = lp[i]
        if k < 0:
            if not (v >= tau) and j < 0:
                j = i
            if v <= -tau:
                k = i
    if k >= 0:
        want[max(j - 1, 0):k + 1] = True
        if j == 0:
            want[S - 1] = True
    else:
        for i in range(S):
            if not (abs(lp[i]) >= tau):
                want[i] = True
                want[S - 1 if i == 0 else i - 1] = True
    if k < 0 or not inside_true:
        want |= ~(lp >= lp.min() + 2.0 * tau)
    return want


CHUNKS = (0, 2, 5, 10, 20, 40, 70, 100)        # chunk_begin() of csrc/tracer.cu


def chunked_screening(lp, tau, inside_true):
    """chunk_gather_kernel / chunk_decide_kernel: the screening pass walks the samples in chunks and stops after the
    first chunk that holds a value <= -tau when the pixel is inside the true mask; what was never evaluated is +inf."""
    seen = np.full(S, np.inf, dtype=np.float32)
    for c in range(len(CHUNKS) - 1):
        sl = slice(CHUNKS[c], CHUNKS[c + 1])
        seen[sl] = lp[sl]
        if inside_true and (lp[sl] <= -tau).any():
            break
    return seen


def candidates_argmin(lp, tau):
    """prefilter_argmin_kernel."""
    return ~(lp >= lp.min() + 2.0 * tau)


def select_sampler(f, inside_true, training):
    """sampler_select_kernel == ray_tracing.py:221-249: returns everything the tracer keeps from the S values."""
    neg = np.nonzero(f < 0)[0]
    zero = np.nonzero(f == 0)[0]
    first = int(neg[0]) if neg.size else (int(zero[0]) if zero.size else S - 1)
    inside_net = bool(f[first] < 0)
    pick = first if (inside_true and inside_net) else int(np.argmin(f))
    sec = (inside_net and inside_true) if training else inside_net
    out = {"first": first, "inside_net": inside_net, "pick": pick, "secant": sec}
    if sec:
        prev = S - 1 if first == 0 else first - 1
        out.update(f_hi=float(f[first]), f_lo=float(f[prev]), prev=prev)
    return out


def _curves(rng, n, tau):
    """Smooth SDF-like profiles along a ray: hits, grazing misses, flat bands inside +-2 tau, exact zeros."""
    t = np.linspace(0.0, 1.0, S)[None, :]
    a = rng.uniform(-0.03, 0.03, (n, 1))
    b = rng.uniform(-0.2, 0.2, (n, 1))
    c = rng.uniform(0.0, 0.5, (n, 1))
    t0 = rng.uniform(-0.2, 1.2, (n, 1))
    f = a + b * (t - t0) + c * (t - t0) ** 2 + 0.002 * np.sin(rng.uniform(5, 40, (n, 1)) * t + rng.uniform(0, 6, (n, 1)))
    f[: n // 10] *= 0.05                                     # whole ray inside the undecidable band
    idx = rng.integers(0, S, n // 20)
    f[np.arange(n // 20) + n // 10, idx] = 0.0               # exact zeros
    f[n // 5: n // 5 + n // 20, 0] = -np.abs(f[n // 5: n // 5 + n // 20, 0]) - 1e-4   # negative first sample: the [-1] wrap
    return f.astype(np.float32)


@pytest.mark.parametrize("tau", [3e-3, 1e-3])
@pytest.mark.parametrize("noise", ["uniform", "plus", "minus", "alternating", "toward_zero"])
def test_sampler_rule_never_changes_the_selection(tau, noise):
    rng = np.random.default_rng(hash((tau, noise)) % (2 ** 32))
    n = 1500
    f = _curves(rng, n, tau)
    e = np.float32(0.999 * tau)
    if noise == "uniform":
        lp = f + rng.uniform(-e, e, f.shape).astype(np.float32)
    elif noise == "plus":
        lp = f + e
    elif noise == "minus":
        lp = f - e
    elif noise == "alternating":
        lp = f + e * np.where(np.arange(S) % 2 == 0, 1.0, -1.0).astype(np.float32)[None, :]
    else:
        lp = f - e * np.sign(f)                               # pushes every sample toward the wrong side of zero
    lp = lp.astype(np.float32)
    assert np.abs(lp.astype(np.float64) - f).max() < tau
    for r in range(n):
        for inside_true, training in ((True, False), (False, True), (True, True)):
            want = candidates_sampler(lp[r], np.float32(tau), inside_true)
            merged = np.where(want, f[r], lp[r])
            assert select_sampler(merged, inside_true, training) == select_sampler(f[r], inside_true, training), (r, noise)
            # same with the chunked screening pass: samples behind the deciding chunk are never evaluated (+inf)
            seen = chunked_screening(lp[r], np.float32(tau), inside_true)
            want = candidates_sampler(seen, np.float32(tau), inside_true)
            merged = np.where(want, f[r], seen)
            assert select_sampler(merged, inside_true, training) == select_sampler(f[r], inside_true, training), (r, noise, "chunked")


@pytest.mark.parametrize("tau", [3e-3, 1e-3])
def test_argmin_rule_never_changes_the_minimum(tau):
    rng = np.random.default_rng(7)
    n = 4000
    f = _curves(rng, n, tau)
    e = np.float32(0.999 * tau)
    for lp in (f + rng.uniform(-e, e, f.shape).astype(np.float32), f - e * np.sign(f - f.mean(axis=1, keepdims=True))):
        lp = lp.astype(np.float32)
        for r in range(n):
            want = candidates_argmin(lp[r], np.float32(tau))
            merged = np.where(want, f[r], lp[r])
            assert int(np.argmin(merged)) == int(np.argmin(f[r]))
            assert merged[np.argmin(merged)] == f[r][np.argmin(f[r])]
