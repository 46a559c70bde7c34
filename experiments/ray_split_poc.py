#!/usr/bin/env python
"""Proof of concept (CPU, exact arithmetic check) for splitting a free-space ray walk into
independent segments WITHOUT changing a single visited voxel -- the round-2 plan for K2
("What comes next").

The reference walk (Octree::computeRayInit / computeRayTakeStep, map/octree.h:1192-1233, driven
by freeSpaceNormal, map/occupancy_map_base.h:1261-1301) accumulates t_max by REPEATED ADDITION and
picks the axis with the smallest t_max, ties x before y before z -- so no closed form may be used.
But the sequence of values t_max[A] takes is a function of axis A alone:
    T_A(1) = t_max0[A],  T_A(i+1) = T_A(i) + t_delta[A]          (i-th A-step happens "at" T_A(i))
and the walk is the merge of the three sequences ordered by (value, axis index).  Hence the state
just before the (a+1)-th step along a chosen axis A is
    count_A = a,                       t_max[A] = T_A(a+1)
    count_B = #{ j : (T_B(j), B) < (T_A(a+1), A) }   for the other two axes, t_max[B] = T_B(count_B + 1)
-- three independent chains of additions and compares, O(steps) cheap operations instead of the
full walk.  Segment 1 is the ordinary walk with its end key replaced by the split key; segment 2
is the ordinary walk started from the split state; it exists iff the split state still satisfies
the loop condition (cur != end and min(t_max) <= distance), because min(t_max) never decreases.

Run: python experiments/ray_split_poc.py   (compares split walks with the full walk on random,
axis-parallel, diagonal (three-way ties) and range-limited rays; uses only float64 numpy scalars).
tests/test_ray_split_poc.py additionally checks walk() against the CPU oracle's free-space set."""
import numpy as np

F = np.float64
DBL_MAX = np.finfo(np.float64).max


def to_key(x, res_factor, max_value):
    return int(np.floor(F(res_factor) * F(x))) + max_value


def key_to_coord(k, res, max_value):
    return (F(np.floor(F(k - max_value))) + F(0.5)) * F(res)


def dda_init(a, b, res, levels):
    """walk from point a towards point b (depth 0): returns None if both are in the same voxel"""
    max_value = 1 << (levels - 1)
    rf = F(1.0) / F(res)
    d = np.array([F(b[i]) - F(a[i]) for i in range(3)], F)
    dist = np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    d = d / dist
    cur = [to_key(a[i], rf, max_value) for i in range(3)]
    end = [to_key(b[i], rf, max_value) for i in range(3)]
    if cur == end:
        return None
    size, hs = F(res), F(res) / F(2.0)
    step, t_delta, t_max = [0] * 3, [F(0)] * 3, [F(0)] * 3
    for i in range(3):
        border = key_to_coord(cur[i], res, max_value) - F(a[i])
        if d[i] > 0:
            step[i], border = 1, border + hs
            t_delta[i], t_max[i] = size / abs(d[i]), border / d[i]
        elif d[i] < 0:
            step[i], border = -1, border - hs
            t_delta[i], t_max[i] = size / abs(d[i]), border / d[i]
        else:
            step[i], t_delta[i], t_max[i] = 0, F(DBL_MAX), F(DBL_MAX)
    return dict(cur=cur, end=end, step=step, t_delta=t_delta, t_max=t_max, dist=dist)


def min_index(t):
    if t[0] <= t[1]:
        return 0 if t[0] <= t[2] else 2
    return 1 if t[1] <= t[2] else 2


def walk(cur, end, step, t_delta, t_max, dist):
    """do { visit; step } while (cur != end && min(t_max) <= dist)  -- freeSpaceNormal's loop"""
    cur, t_max, out = list(cur), list(t_max), []
    while True:
        out.append(tuple(cur))
        i = min_index(t_max)
        cur[i] += step[i]
        t_max[i] = t_max[i] + t_delta[i]
        if cur == list(end) or not (min(min(t_max[0], t_max[1]), t_max[2]) <= dist):
            break
    return out


def split_state(s, axis, a):
    """state just before the (a+1)-th step along `axis`, from three independent addition chains"""
    T = s["t_max"][axis]
    for _ in range(a):                       # T_A(a + 1)
        T = T + s["t_delta"][axis]
    cur, t_max = list(s["cur"]), list(s["t_max"])
    cur[axis] += a * s["step"][axis]
    t_max[axis] = T
    for b in range(3):
        if b == axis:
            continue
        t, cnt = s["t_max"][b], 0
        # steps of axis b that come before: (t, b) < (T, axis)
        while (t < T) or (t == T and b < axis):
            if s["step"][b] == 0:
                break
            t = t + s["t_delta"][b]
            cnt += 1
        cur[b] += cnt * s["step"][b]
        t_max[b] = t
    return cur, t_max


def split_walk(s, parts=2):
    """the same voxel sequence from `parts` independent segments"""
    span = [abs(s["end"][i] - s["cur"][i]) for i in range(3)]
    axis = int(np.argmax(span))
    cuts = [span[axis] * k // parts for k in range(1, parts)]
    states = [(list(s["cur"]), list(s["t_max"]))]
    for a in cuts:
        if a == 0:
            continue
        cur, t_max = split_state(s, axis, a)
        # the segment exists iff the full walk gets there: loop condition still true in that state
        if cur == s["end"] or not (min(min(t_max[0], t_max[1]), t_max[2]) <= s["dist"]):
            break
        if cur == states[-1][0]:
            continue
        states.append((cur, t_max))
    out = []
    for k, (cur, t_max) in enumerate(states):
        seg_end = states[k + 1][0] if k + 1 < len(states) else s["end"]
        out.append(walk(cur, seg_end, s["step"], s["t_delta"], t_max, s["dist"]))
    return out


def main():
    rng = np.random.default_rng(5)
    res, levels = 0.02, 16
    total = segs = 0
    for case in range(3000):
        a = rng.uniform(-3, 3, 3)
        b = a + rng.normal(size=3) * rng.uniform(0.05, 8.0)
        kind = case % 6
        if kind == 1:                                  # axis parallel
            b = a.copy()
            b[rng.integers(0, 3)] += rng.uniform(-6, 6)
        elif kind == 2:                                # exact diagonal from a voxel corner: 3-way ties
            a = np.round(a / res) * res
            b = a + np.array([1, 1, 1]) * res * int(rng.integers(3, 150)) * rng.choice([-1, 1])
        elif kind == 3:                                # two equal components from a voxel centre
            a = (np.round(a / res) + 0.5) * res
            n = int(rng.integers(3, 100))
            b = a + np.array([n, n, rng.integers(-50, 50)]) * res
        a32, b32 = a.astype(np.float32).astype(np.float64), b.astype(np.float32).astype(np.float64)
        s = dda_init(a32, b32, res, levels)
        if s is None:
            continue
        if kind == 4:                                  # range-limited walk: stops on min(t_max) > dist
            s["dist"] = s["dist"] * F(rng.uniform(0.2, 0.9))
        full = walk(s["cur"], s["end"], s["step"], s["t_delta"], s["t_max"], s["dist"])
        for parts in (2, 3, 4, 7):
            pieces = split_walk(s, parts)
            joined = [v for p in pieces for v in p]
            assert joined == full, (case, parts, len(joined), len(full))
            segs += len(pieces)
        total += 1
    print("rays checked:", total, "| segments:", segs, "| every split walk == the full walk, voxel for voxel")


if __name__ == "__main__":
    main()
