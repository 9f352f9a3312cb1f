"""A NumPy model of the direct tile binning (binocular3dgs_b200/csrc/binning.cu, tile_bins_kernel and the
three scan kernels), run on the CPU: the arithmetic the CUDA kernels rely on — per-warp private counters as
ranks, per-warp counts packed one byte each, the table scanned along the batches and over the tiles, block-
local tile-major staging slots with a capacity and direct stores beyond it, the packed `lane / width`
reciprocal — must reproduce what the reference gets from one stable sort of (tile, depth) keys
(rasterizer_impl.cu:70-138, :304-309).  Test infrastructure: the product path is the CUDA kernel, which the
`-m gpu` tests compare with the reference itself; this model pins the ALGORITHM, with small batch / band /
staging sizes so that every corner (band clipping, rectangles wider than a warp, staging overflow, empty
warps, ragged last batch) is hit on a few hundred rectangles.
"""
import numpy as np
import pytest

WARPS = 8


def reference_lists(rects, order, grid_x, grid_y):
    """Stable sort by tile of the instance stream emitted in depth order."""
    tiles, ids = [], []
    for g in order:
        x0, y0, w, h = rects[g]
        for y in range(y0, y0 + h):
            for x in range(x0, x0 + w):
                tiles.append(y * grid_x + x)
                ids.append(g)
    tiles, ids = np.array(tiles, dtype=np.int64), np.array(ids, dtype=np.int64)
    perm = np.argsort(tiles, kind="stable")
    T = grid_x * grid_y
    counts = np.bincount(tiles, minlength=T)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    ranges = np.stack([np.where(counts > 0, starts, 0), np.where(counts > 0, starts + counts, 0)], axis=1)
    return ids[perm], ranges


def lane_tiles(t0, bw, bh, grid_x):
    """Tiles of one clipped rectangle in the order the warp's steps visit them, as (step, tile) —
    lane L handles column L % bw of rows L / bw, L / bw + rows, ... with rows = 32 / bw (binning.cu)."""
    out = []
    if bw <= 32:
        inv = 1024 // bw + 1                         # the packed reciprocal: (L * inv) >> 10 == L // bw for L <= 32
        rows = (32 * inv) >> 10
        for lane in range(32):
            ry = (lane * inv) >> 10
            assert ry == lane // bw and rows == 32 // bw
            if ry >= rows:
                continue
            y, step = ry, 0
            while y < bh:
                out.append((step, t0 + (lane - ry * bw) + y * grid_x))
                y += rows
                step += 1
    else:
        step = 0
        for y in range(bh):
            for xb in range(0, bw, 32):
                for lane in range(min(32, bw - xb)):
                    out.append((step, t0 + y * grid_x + xb + lane))
                step += 1
    return out


def model(rects, order, grid_x, grid_y, B, band_rows, stage_cap):
    P, T = len(order), grid_x * grid_y
    nb = (P + B - 1) // B
    sub = B // WARPS
    assert sub <= 255
    table = np.zeros((nb, T), dtype=np.int64)
    wcount = np.zeros((nb, T, WARPS), dtype=np.uint8)

    def walk(b, by0, by1, counters, visit):
        """One band of one batch: warp by warp (the order between warps is irrelevant: private counters)."""
        n = min(P, (b + 1) * B) - b * B
        for w in range(WARPS):
            for k in range(w * sub, min(n, (w + 1) * sub)):
                g = order[b * B + k]
                x0, y0, bw, h = rects[g]
                cy0, cy1 = max(y0, by0), min(y0 + h, by1)
                if bw == 0 or h == 0 or cy1 <= cy0:
                    continue
                t0 = (cy0 - by0) * grid_x + x0
                seen = set()
                for _, t in lane_tiles(t0, bw, cy1 - cy0, grid_x):
                    assert t not in seen             # the lanes of one Gaussian touch distinct counters
                    seen.add(t)
                    visit(w, k, g, t, counters[w][t])
                    counters[w][t] += 1

    bands = [(by0, min(grid_y, by0 + band_rows)) for by0 in range(0, grid_y, band_rows)]
    # count pass
    for b in range(nb):
        for by0, by1 in bands:
            tiles = (by1 - by0) * grid_x
            counters = np.zeros((WARPS, tiles), dtype=np.int64)
            walk(b, by0, by1, counters, lambda *a: None)
            assert counters.max(initial=0) <= 255
            wcount[b, by0 * grid_x:by1 * grid_x, :] = counters.T.astype(np.uint8)
            table[b, by0 * grid_x:by1 * grid_x] = counters.sum(axis=0)
    # scans: along the batches per tile, over the tiles
    totals = table.sum(axis=0)
    tile_start = np.concatenate([[0], np.cumsum(totals)[:-1]])
    ranges = np.stack([np.where(totals > 0, tile_start, 0), np.where(totals > 0, tile_start + totals, 0)], axis=1)
    scanned = tile_start[None, :] + np.cumsum(table, axis=0) - table
    # scatter pass
    out = np.full(int(totals.sum()), -1, dtype=np.int64)
    staged_total = direct_total = 0
    for b in range(nb):
        for by0, by1 in bands:
            sl = slice(by0 * grid_x, by1 * grid_x)
            wc = wcount[b, sl, :].astype(np.int64)                    # [tiles][warps]
            block = wc.sum(axis=1)
            # the kernel sums the eight bytes with two masked adds per word: same value
            packed = wc[:, 0] | wc[:, 1] << 8 | wc[:, 2] << 16 | wc[:, 3] << 24, wc[:, 4] | wc[:, 5] << 8 | wc[:, 6] << 16 | wc[:, 7] << 24
            v = sum((p & 0x00ff00ff) + ((p >> 8) & 0x00ff00ff) for p in packed)
            assert ((v & 0xffff) + (v >> 16) == block).all()
            local = np.concatenate([[0], np.cumsum(block)[:-1]])      # slot of a tile's first instance in the block's buffer
            n_local = int(block.sum())
            assert n_local <= 65535                                   # else the kernel takes its 32-bit turn-taking path
            base = scanned[b, sl] - local                             # s_base: global position - local slot
            counters = (local[:, None] + np.cumsum(wc, axis=1) - wc).T.copy()   # [warps][tiles]: the warps' prefixes on top
            stage = {}

            def visit(w, k, g, t, slot):
                nonlocal staged_total, direct_total
                if slot < stage_cap:
                    assert slot not in stage
                    stage[slot] = (k, t)                              # one word: index in the batch << 16 | tile
                    staged_total += 1
                else:
                    assert out[base[t] + slot] == -1
                    out[base[t] + slot] = g
                    direct_total += 1

            walk(b, by0, by1, counters, visit)
            for slot, (k, t) in stage.items():                        # copy-out: slot i goes to s_base[tile] + i
                pos = base[t] + slot
                assert out[pos] == -1
                out[pos] = order[b * B + k]
    return out, ranges, staged_total, direct_total


@pytest.mark.parametrize("seed,grid_x,grid_y,P,B,band_rows,stage_cap", [
    (0, 12, 9, 300, 64, 2, 40),        # several bands, staging overflows often
    (1, 40, 6, 250, 64, 3, 10 ** 6),   # rectangles wider than a warp, everything staged
    (2, 7, 23, 333, 128, 5, 0),        # ragged last batch, nothing staged (the direct-store variant)
    (3, 33, 4, 90, 256, 4, 64),        # one batch, one band, empty warps
])
def test_model_equals_stable_sort(seed, grid_x, grid_y, P, B, band_rows, stage_cap):
    rng = np.random.default_rng(seed)
    rects = []
    for _ in range(P):
        if rng.random() < 0.25:
            rects.append((0, 0, 0, 0))                               # culled: no tiles
            continue
        w = int(min(grid_x, rng.choice([1, 2, 3, 5, 8, 13, 33, 40]))) if rng.random() < 0.8 else grid_x
        h = int(min(grid_y, rng.integers(1, 7)))
        rects.append((int(rng.integers(0, grid_x - w + 1)), int(rng.integers(0, grid_y - h + 1)), w, h))
    order = rng.permutation(P)                                        # depth order
    want_list, want_ranges = reference_lists(rects, order, grid_x, grid_y)
    got_list, got_ranges, staged, direct = model(rects, order, grid_x, grid_y, B, band_rows, stage_cap)
    assert (got_list == want_list).all()
    assert (got_ranges == want_ranges).all()
    if stage_cap == 0:
        assert staged == 0
    elif stage_cap >= 10 ** 6:
        assert direct == 0
    else:
        assert staged > 0 and direct > 0


def test_packed_reciprocal_is_exact():
    """(L * (1024 // w + 1)) >> 10 == L // w for every lane count L <= 32 and width w <= 32, and the
    reciprocal fits the 11 bits it is packed in next to an 11-bit width and an 8-bit height."""
    for w in range(1, 33):
        inv = 1024 // w + 1
        assert inv < (1 << 11)
        for L in range(33):
            assert (L * inv) >> 10 == L // w
    assert 2040 < (1 << 11) and 200 < 255
