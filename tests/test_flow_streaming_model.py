"""CPU model of the arithmetic of the streaming Farneback iteration kernel (k4_flow_iter_march, flow_kernels.cu) against the
oracle's exact (f64) 15 x 15 box sums + solve (oracle/farneback.py:box_solve, OpenCV's FarnebackUpdateFlow_Blur).

The kernel keeps, per column, running sums over a whole row segment (hundreds of rows) instead of summing 15 values afresh,
so the question the GPU tests answer only on a few clips is answered here on adversarial inputs: a column whose
structure-tensor entries drop by seven orders of magnitude (a strong edge above a flat, near-singular region) must not
leave rounding residue in the sums below it.  Three accumulators are modelled: f64 (the kernel's default, and what OpenCV
itself uses), Kahan-compensated fp32 (flow_impl 1) and the plain fp32 sliding sum of the earlier tile kernel (flow_impl 2);
only the first survives that input, which is why it is the default."""
import numpy as np
import pytest

from oracle import farneback as FB

F32 = np.float32


def _vertical_running_sums(M, rows_per_seg, mode):
    """Per segment, rows ya-7 .. yb+6 (replicate border) are added one by one; from the 16th row on the row that left the
    15-row window is subtracted.  mode: "f64" (sums in double, rounded to fp32 per output row), "kahan" (fp32, Kahan
    summation of the fp32 increments), "plain" (fp32)."""
    h, w, _ = M.shape
    out = np.empty_like(M)
    for ya in range(0, h, rows_per_seg):
        yb = min(ya + rows_per_seg, h)
        vs = np.zeros((w, 5), np.float64 if mode == "f64" else F32)
        comp = np.zeros((w, 5), F32)
        ring = []
        for k in range(yb - ya + 14):
            m = M[min(max(ya - 7 + k, 0), h - 1)]
            if mode == "f64":
                inc = m.astype(np.float64) - ring[k - 15].astype(np.float64) if k >= 15 else m.astype(np.float64)
                ring.append(m)
                vs = vs + inc
                if k >= 14:
                    out[ya + k - 14] = vs.astype(F32)
                continue
            inc = m - ring[k - 15] if k >= 15 else m
            ring.append(m)
            if mode == "kahan":
                yk = (inc - comp).astype(F32)
                t = (vs + yk).astype(F32)
                comp = ((t - vs).astype(F32) - yk).astype(F32)
                vs = t
            else:
                vs = (vs + inc).astype(F32)
            if k >= 14:
                out[ya + k - 14] = vs
    return out


def _horizontal_and_solve(V):
    """15-column sums of the vertical sums (fp32, sequential, replicate border) + the f64 2 x 2 solve of the kernel."""
    h, w, _ = V.shape
    xs = np.clip(np.arange(-7, w + 7), 0, w - 1)
    P = V[:, xs]
    S = np.zeros((h, w, 5), F32)
    for j in range(15):
        S = (S + P[:, j:j + w]).astype(F32)
    S = S.astype(np.float64) * (1.0 / 225.0)
    g11, g12, g22, h1, h2 = (S[..., i] for i in range(5))
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    return np.stack([((g11 * h2 - g12 * h1) * idet), ((g22 * h1 - g12 * h2) * idet)], -1).astype(F32)


def _structure_tensor(h, w, seed, edge_rows=None):
    """M = (r4^2+r6^2, (r4+r5) r6, r5^2+r6^2, r4 r2+r6 r3, r6 r2+r5 r3) from random polynomial coefficients, optionally with a
    band of rows whose coefficients are 3000 x larger (M 1e7 x larger) than the rest."""
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((h, w, 5)).astype(F32) * F32(0.02)
    if edge_rows is not None:
        r[edge_rows[0]:edge_rows[1]] *= F32(3000.0)
    r2, r3, r4, r5, r6 = (r[..., i] for i in range(5))
    return np.stack([r4 * r4 + r6 * r6, (r4 + r5) * r6, r5 * r5 + r6 * r6, r4 * r2 + r6 * r3, r6 * r2 + r5 * r3], -1).astype(F32)


@pytest.mark.parametrize("mode", ["f64", "kahan"])
@pytest.mark.parametrize("rows_per_seg", [216, 48, 8])
def test_streaming_sums_match_exact_box_solve(rows_per_seg, mode):
    M = _structure_tensor(240, 64, seed=1)
    ref = FB.box_solve(M)
    got = _horizontal_and_solve(_vertical_running_sums(M, rows_per_seg, mode))
    assert np.abs(got - ref).max() < 2e-5 * max(1.0, float(np.abs(ref).max()))


def test_f64_sums_survive_a_strong_edge_above_a_flat_region():
    M = _structure_tensor(200, 48, seed=2, edge_rows=(20, 40))
    ref = FB.box_solve(M)
    flat = slice(70, 200)                      # rows whose 15-row window no longer touches the edge band
    scale = float(np.abs(ref[flat]).max())
    err = {mode: float(np.abs(_horizontal_and_solve(_vertical_running_sums(M, 216, mode))[flat] - ref[flat]).max())
           for mode in ("f64", "kahan", "plain")}
    assert err["f64"] < 1e-4 * scale, (err, scale)
    # fp32 walks keep the edge's rounding residue (the increment m - old is already rounded), compensated or not
    assert err["kahan"] > 100 * err["f64"] and err["plain"] > 100 * err["f64"], err


def test_segment_split_changes_only_the_last_bits():
    M = _structure_tensor(120, 40, seed=3)
    a = _horizontal_and_solve(_vertical_running_sums(M, 216, "f64"))
    b = _horizontal_and_solve(_vertical_running_sums(M, 24, "f64"))
    assert np.abs(a - b).max() < 1e-5 * max(1.0, float(np.abs(a).max()))
