"""How much would an exact bound hierarchy prune in the loop-closure search? (CPU study for the next round; no GPU.)

For a few (node, submap) pairs of the loop-closure bench's shape it scores every leaf of the stock window densely in numpy,
then evaluates the reference's own bounds (sliding-window maxima of the 8-bit grid, precomputation_grid_3d.cc:62-81) at block
sizes 8, 4 and 2 and counts which blocks a best-first search with threshold max(min_score, best leaf) still has to open.
"""
import os
import sys

import numpy as np
from scipy.ndimage import maximum_filter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]


def main():
    import orc
    from helpers import apply_pose, workload
    w = workload(beams=64, num_map_scans=20, num_scans=3)
    xs, ys, zs, vs = w["hi"].export()
    f = np.float32
    prob = np.array([orc.lib().orc_value_to_probability(int(v)) for v in np.unique(vs)], f)
    lut = dict(zip(np.unique(vs).tolist(), np.rint((prob - f(0.1)) * (f(255.0) / f(0.8))).astype(np.int64).tolist()))
    v8 = np.array([lut[int(v)] for v in vs], np.uint8)
    lo3 = np.array([xs.min(), ys.min(), zs.min()]) - 200
    shape = np.array([xs.max(), ys.max(), zs.max()]) + 200 - lo3 + 1
    vol = np.zeros(shape, np.uint8)
    vol[xs - lo3[0], ys - lo3[1], zs - lo3[2]] = v8
    wxy, wz, res = 50, 10, 0.1
    rng = np.random.default_rng(5)
    min_score = 0.3
    for k in range(5):
        unmatched = k >= 3     # last cases: the node is 12 m (across the street) / 6 m (along it) away from where it really was
        kk = min(k, 2)
        shift = {3: [0.0, 12.0, 0.0], 4: [6.0, 0.0, 0.0]}.get(k, [0.0, 0.0, 0.0])
        pts = orc.ingest_scan(w["opts"], w["scans"][kk], w["origin"], w["prev"][kk], w["truth"][kk])["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        guess = np.array(w["truth"][kk], np.float64)
        guess[:3] += rng.uniform(-1, 1, 3) * [2.0, 2.0, 0.4] + np.array(shift)
        cells = np.array([w["hi"].cell_index(p) for p in apply_pose(guess, pts[hk].astype(np.float64)).astype(np.float32)]) - lo3
        n = len(cells)
        sums = np.zeros((2 * wxy + 1, 2 * wxy + 1, 2 * wz + 1), np.int64)
        for c in cells:
            sums += vol[c[0] - wxy:c[0] + wxy + 1, c[1] - wxy:c[1] + wxy + 1, c[2] - wz:c[2] + wz + 1]
        best = sums.max()
        floor = (min_score - 0.1) / 0.8 * 255 * n     # a leaf needs sum > floor to exceed min_score
        thr = max(best, floor) if best > floor else floor
        line = [f"pair {k}{' (unmatched)' if unmatched else ''}: N={n} best leaf sum/N = {best / n:.1f} (score {0.1 + best / n * 0.8 / 255:.3f}); "
                f"threshold = {'best leaf' if best > floor else 'min_score %.2f' % min_score}; leaves {sums.size}"]
        for b in (8, 4, 2):
            # bound of the block of offsets [o, o+b)^3 = sum_i max over that window of vol around c_i  (exact sliding max)
            m = maximum_filter(vol, size=b, origin=-(b // 2) if b % 2 == 0 else 0, mode="constant")   # window [x, x+b)
            starts = [np.arange(-wxy, wxy + 1, b), np.arange(-wxy, wxy + 1, b), np.arange(-wz, wz + 1, b)]
            bounds = np.zeros((len(starts[0]), len(starts[1]), len(starts[2])), np.int64)
            for c in cells:
                bounds += m[np.ix_(c[0] + starts[0], c[1] + starts[1], c[2] + starts[2])]
            # sanity: every block bound dominates its leaves
            chk = sums[:b * (sums.shape[0] // b), :b * (sums.shape[1] // b), :b * (sums.shape[2] // b)]
            blk = chk.reshape(chk.shape[0] // b, b, chk.shape[1] // b, b, chk.shape[2] // b, b).max(axis=(1, 3, 5))
            assert np.all(bounds[:blk.shape[0], :blk.shape[1], :blk.shape[2]] >= blk)
            if b == 8:
                top = bounds.max()
                line.append(f"  8^3 bounds: max {top / n:.1f}, blocks within 1/8 of it: {int((bounds >= top - top // 8).sum())}, within 1/64: "
                            f"{int((bounds >= top - top // 64).sum())}, equal to it: {int((bounds == top).sum())}; best leaf inside the top block: "
                            f"{blk[np.unravel_index(np.argmax(bounds[:blk.shape[0], :blk.shape[1], :blk.shape[2]]), blk.shape)] / n:.1f}")
            open_blocks = int((bounds >= thr).sum())
            line.append(f"  block {b}^3: {bounds.size} blocks, {open_blocks} with bound >= threshold ({100.0 * open_blocks / bounds.size:.1f}%)"
                        f" -> {open_blocks * b ** 3} leaves ({100.0 * open_blocks * b ** 3 / sums.size:.1f}% of brute force)")
        print("\n".join(line))


if __name__ == "__main__":
    main()
