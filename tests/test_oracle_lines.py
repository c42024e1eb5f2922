"""The line-extraction oracle (oracle/laser_lines.hpp = laser_manager::spawn_scan + scan::add_line, reference
src/trajectory/laser_manager.cpp:350-422, :137-154) pinned on the CPU.  The reference ships no tests or fixtures for
this path ("parity unpinned"), so the pins are: numpy.linalg.svd for the JacobiSVD restatement, an independent numpy
re-derivation of the whole step (tests/independent_lines.py), hand-built scans with known answers, and the frozen
golden vectors in tests/golden/lines_scans.npz."""
import os

import numpy as np
import pytest

import lvio2d_b200 as L
from independent_lines import spawn_scan

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "lines_scans.npz")


@pytest.fixture(scope="module")
def lp():
    return L.corridor_line_params()


def wall(p, q, n, noise=0.0, rng=None):
    t = np.linspace(0.0, 1.0, n)[:, None]
    pts = (1 - t) * np.asarray(p, float) + t * np.asarray(q, float)
    if noise:
        pts = pts + rng.normal(0.0, noise, pts.shape)
    return pts


def test_fit_matches_numpy_svd(oracle):
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(3, 300))
        a, p0 = rng.uniform(0, 2 * np.pi), rng.uniform(-20, 20, 2)
        s = np.sort(rng.uniform(0, rng.uniform(0.1, 8.0), n))
        pts = p0 + np.outer(s, [np.cos(a), np.sin(a)]) + rng.normal(0, rng.choice([0.0, 1e-3, 1e-2]), (n, 2))
        v = oracle.fit_line(pts)
        w = np.linalg.svd(np.c_[pts, np.ones(n)], full_matrices=False)[2][2]
        w = w * np.sign(w[np.argmax(np.abs(w))])
        assert np.abs(v - w).max() < 1e-9


def test_known_answers(oracle, lp):
    # an L-shaped corner: two perfect walls of 40 points each -> two lines split at the corner point
    # (rotated by 0.3 rad: create_line's `fabs(b) < 0.5 // b==0` test acts on the 3-D-normalised (a, b, c), so an exactly
    # axis-aligned line away from the origin divides by a ~ 0 in the reference — a quirk the oracle reproduces)
    R = np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
    A, B, Cc = R @ [1, -2], R @ [1, 2], R @ [-3, 2]
    a, b = wall(A, B, 40), wall(B, Cc, 40)[1:]
    pts = np.concatenate([a, b])
    n, lines, abc, rng = oracle.extract_lines(lp, [0, len(pts)], pts)
    assert n[0] == 2
    assert rng[0, 0].tolist() == [0, 39] and rng[0, 1].tolist() == [39, 78]
    assert np.allclose(lines[0, 0], np.r_[A, B], atol=1e-9) and np.allclose(lines[0, 1], np.r_[B, Cc], atol=1e-9)
    # a gap wider than line_continuous_threshold splits; a 2-point tail is dropped (needs 3 points)
    pts = np.concatenate([wall([2, 0.1], [2.2, 1], 20), wall([2.4, 2], [2.6, 3], 20), [[5.0, 5.0], [5.0, 5.2]]])
    n, lines, _, rng = oracle.extract_lines(lp, [0, len(pts)], pts)
    assert n[0] == 2 and rng[0, :2].tolist() == [[0, 19], [20, 39]]
    # empty scan, and a scan too short for any line
    n, _, _, _ = oracle.extract_lines(lp, [0, 0, 2], np.array([[1.0, 0.0], [1.0, 0.1]]))
    assert n.tolist() == [0, 0]
    # a line whose points all fall outside the w x h grid never enters scan::lines
    n, _, _, _ = oracle.extract_lines(lp, [0, 30], wall([60, 0], [60.5, 3], 30))
    assert n[0] == 0
    # the residual filter: a noisy blob wider than line_max_dis gives no line over its whole span
    g = np.random.default_rng(0)
    blob = wall([3, 0], [3, 0.3], 30) + g.uniform(-0.3, 0.3, (30, 2))
    n, lines, _, rng = oracle.extract_lines(lp, [0, 30], blob)
    for k in range(n[0]):
        seg = blob[rng[0, k, 0]:rng[0, k, 1] + 1]
        d = lines[0, k, 2:] - lines[0, k, :2]
        u = d / np.linalg.norm(d)
        rel = seg - lines[0, k, :2]
        assert np.abs(rel[:, 0] * u[1] - rel[:, 1] * u[0]).max() <= lp.line_max_dis + 1e-12


@pytest.mark.parametrize("sigma", [0.002, 0.005, 0.01, 0.03])
def test_oracle_matches_independent_restatement(oracle, lp, sigma):
    """(Noisy scans only: on noise-free walls every corner response is -1 up to rounding, the non-maximum suppression
    then ranks rounding errors, and two correct implementations may legitimately pick different candidates.)"""
    off, pts = L.synth.make_scan_batch(6, 7, range_sigma=sigma)
    n, lines, abc, rng = oracle.extract_lines(lp, off, pts)
    for s in range(len(off) - 1):
        want = spawn_scan(lp, pts[off[s]:off[s + 1]])
        assert n[s] == len(want)
        for k, (i1, i2, e1, e2, v) in enumerate(want):
            assert rng[s, k].tolist() == [i1, i2]
            assert np.abs(lines[s, k] - np.r_[e1, e2]).max() < 1e-8
            assert np.abs(abc[s, k] - v).max() < 1e-8


def test_properties(oracle, lp):
    off, pts = L.synth.make_scan_batch(16, 11)
    n, lines, abc, rng = oracle.extract_lines(lp, off, pts)
    assert n.max() <= lines.shape[1]
    for s in range(len(off) - 1):
        p = pts[off[s]:off[s + 1]]
        r = rng[s, :n[s]]
        assert np.all(r[:, 1] - r[:, 0] >= 2) and np.all(r[1:, 0] >= r[:-1, 1])      # ordered, share at most end points
        for k in range(n[s]):
            a, b, c = abc[s, k]
            e = lines[s, k].reshape(2, 2)
            assert np.abs(e @ [a, b] + c).max() / np.hypot(a, b) < 1e-9                 # end points lie on the fitted line
            seg = p[r[k, 0]:r[k, 1] + 1]
            assert (np.abs(seg @ [a, b] + c) / np.hypot(a, b)).max() <= lp.line_max_dis + 1e-9
            assert np.linalg.norm(e[0] - e[1]) >= lp.line_min_len


def test_golden_vectors(oracle, lp):
    """Frozen inputs + outputs (scripts/make_golden.py --lines): the oracle must keep producing them."""
    G = np.load(GOLDEN)
    n, lines, abc, rng = oracle.extract_lines(lp, G["point_offset"], G["points"], max_lines=G["lines"].shape[1])
    assert np.array_equal(n, G["n_lines"]) and np.array_equal(rng, G["index_range"])
    assert np.abs(lines - G["lines"]).max() < 1e-9 and np.abs(abc - G["abc"]).max() < 1e-9
