"""Synthetic inputs of the shapes BASELINE.json names (descriptor sets, correspondence scenes, image grids).

Pure numpy, seeded; used by bench.py and the tests. Nothing here touches the oracle or the reference.
"""
import numpy as np

DESCRIPTOR_BITS = 486
ROW_WORDS = 8


def random_descriptors(n, rng):
    """n rows of 486 iid Bernoulli(1/2) bits in the 64-byte std::bitset<486> image (bits 486..511 zero)."""
    rows = rng.integers(0, 1 << 63, size=(n, ROW_WORDS), dtype=np.uint64) << np.uint64(1)
    rows |= rng.integers(0, 2, size=(n, ROW_WORDS), dtype=np.uint64)
    rows[:, 7] &= np.uint64((1 << (DESCRIPTOR_BITS - 7 * 64)) - 1)
    return rows


def flip_bits(rows, p, rng):
    """Flip each of the 486 descriptor bits with probability p."""
    bits = np.unpackbits(np.ascontiguousarray(rows).view(np.uint8).reshape(len(rows), 64), axis=1, bitorder="little")
    flips = (rng.random((len(rows), 512)) < p).astype(np.uint8)
    flips[:, DESCRIPTOR_BITS:] = 0
    bits ^= flips
    return np.packbits(bits, axis=1, bitorder="little").view(np.uint64).reshape(len(rows), ROW_WORDS).copy()


def config2_pair(n1=10000, n2=10000, seed=1, match_fraction=0.5, noise=0.08):
    """set A random; set B = noisy copies of a permutation of A for the first half (so the ratio
    test passes for a controlled subset and there are genuine ties), fresh random rows for the rest."""
    rng = np.random.default_rng(seed)
    a = random_descriptors(n1, rng)
    b = random_descriptors(n2, rng)
    m = int(min(n1, n2) * match_fraction)
    perm = rng.permutation(n1)[:m]
    b[:m] = flip_bits(a[perm], noise, rng)
    return a, b


def homography_scene(n_inliers, n_outliers, seed=42, noise=0.0):
    """Correspondences [n][7] (unit-depth rays + quality 0) under a fixed ground-truth homography."""
    rng = np.random.default_rng(seed)
    c, s = np.cos(0.1), np.sin(0.1)
    H = np.array([[c, -s, 0.005], [s, c, -0.003], [0, 0, 1.0]])
    p1 = np.concatenate([rng.uniform(-1, 1, (n_inliers, 2)), np.ones((n_inliers, 1))], axis=1)
    p2 = p1 @ H.T
    p2 /= p2[:, 2:3]
    if noise:
        p2[:, :2] += rng.normal(0, noise, (n_inliers, 2))
    o1 = np.concatenate([rng.uniform(-2, 2, (n_outliers, 2)), np.ones((n_outliers, 1))], axis=1)
    o2 = np.concatenate([rng.uniform(-2, 2, (n_outliers, 2)), np.ones((n_outliers, 1))], axis=1)
    corr = np.zeros((n_inliers + n_outliers, 7))
    corr[:n_inliers, 0:3], corr[:n_inliers, 3:6] = p1, p2
    corr[n_inliers:, 0:3], corr[n_inliers:, 3:6] = o1, o2
    return corr, H


def random_models(kind, h, seed=3, base=None):
    """h random 3x3 models as [h][18] (matrix column-major + inverse for the homography)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((h, 18))
    for i in range(h):
        if kind == 0:
            M = (np.eye(3) if base is None else base) + rng.normal(0, 0.02, (3, 3))
            M /= M[2, 2]
            out[i, :9] = M.T.ravel()
            out[i, 9:] = np.linalg.inv(M).T.ravel()
        else:
            M = rng.normal(0, 1, (3, 3))
            out[i, :9] = M.T.ravel()
    return out


def grid_survey(rows, cols, n_desc, seed=7, k_nn=10, noise=0.08, overlap=0.6):
    """a rows x cols grid of camera positions; every image holds n_desc descriptors drawn from a
    shared pool of world points by footprint overlap (+ bit noise) so neighbours genuinely match; directed pairs =
    k_nn nearest in position minus self (mirrors src/pipeline/link_stage.cpp:26-34). Returns (list of descriptor
    arrays, positions, pairs)."""
    rng = np.random.default_rng(seed)
    n_img = rows * cols
    pos = np.stack(np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64)), -1).reshape(-1, 2)
    # world points live on a lattice of cells; an image sees the cells within its footprint
    per_cell = max(1, int(n_desc * (1 - overlap) ** 2 / 1.0))
    foot = 1.0 / (1 - overlap)  # footprint width in grid units
    half = foot / 2
    cell_cache = {}

    def cell_points(cx, cy):
        key = (cx, cy)
        if key not in cell_cache:
            r = np.random.default_rng([seed, cx + 1000, cy + 1000])
            cell_cache[key] = random_descriptors(per_cell, r)
        return cell_cache[key]

    images = []
    for i in range(n_img):
        x, y = pos[i]
        cells = [(cx, cy) for cx in range(int(np.floor(x - half)), int(np.ceil(x + half)))
                 for cy in range(int(np.floor(y - half)), int(np.ceil(y + half)))]
        pool = np.concatenate([cell_points(cx, cy) for cx, cy in cells])
        if len(pool) >= n_desc:
            sel = rng.permutation(len(pool))[:n_desc]
            d = pool[sel]
        else:
            d = np.concatenate([pool, random_descriptors(n_desc - len(pool), rng)])
        images.append(flip_bits(d, noise, rng))
    pairs = []
    for i in range(n_img):
        d2 = ((pos - pos[i]) ** 2).sum(1)
        nn = np.argsort(d2, kind="stable")[:k_nn]
        for j in nn:
            if j != i:
                pairs.append((i, int(j)))
    return images, pos, pairs


class PlanarSurvey:
    """SURVEY 8d c4/c5 with geometry: a rows x cols grid of nadir cameras over a textured plane. World points live
    in unit cells of the ground plane (`per_cell` points each, position and descriptor derived from (seed, cell), so
    any image can be generated on its own, on any rank); an image sees the points inside its footprint
    (`footprint` grid units wide) at pixel = principal point + (world - camera position) * pixels_per_unit + noise,
    with `noise` descriptor bit flips; it is padded with clutter up to n_desc features. Directed pairs = k_nn nearest
    cameras minus self (mirrors src/pipeline/link_stage.cpp:26-34). Feature order = strength descending, like the
    reference's extractor leaves it."""

    def __init__(self, rows, cols, n_desc, seed=7, k_nn=10, noise=0.08, footprint=2.5, image_size=(4000, 3000),
                 focal=3000.0, pixel_noise=0.3):
        self.rows, self.cols, self.n_desc, self.seed, self.noise = rows, cols, n_desc, seed, noise
        self.footprint, self.image_size, self.focal, self.pixel_noise = footprint, image_size, focal, pixel_noise
        self.n_images = rows * cols
        self.positions = np.stack(np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64)),
                                  -1).reshape(-1, 2)
        # ~70 % of an image's features are world points, the rest clutter
        self.per_cell = max(1, int(0.7 * n_desc / (footprint * footprint * image_size[1] / image_size[0])))
        self.pairs = []
        for i in range(self.n_images):
            d2 = ((self.positions - self.positions[i]) ** 2).sum(1)
            for j in np.argsort(d2, kind="stable")[:k_nn]:
                if j != i:
                    self.pairs.append((i, int(j)))
        self._cells = {}

    def camera8(self):
        w, h = self.image_size
        return np.array([self.focal, w / 2, h / 2, 0, 0, 0, 0, 0], np.float64)

    def _cell(self, cx, cy):
        key = (cx, cy)
        if key not in self._cells:
            r = np.random.default_rng([self.seed, cx + 100000, cy + 100000])
            xy = r.random((self.per_cell, 2)) + np.array([cx, cy], np.float64)
            self._cells[key] = (xy, random_descriptors(self.per_cell, r))
        return self._cells[key]

    def image(self, i):
        """-> (descriptors [n][8] u64, xy [n][2] pixels, strength [n] f32)"""
        rng = np.random.default_rng([self.seed, 7919, i])
        w, h = self.image_size
        px_per_unit = w / self.footprint
        half = np.array([self.footprint / 2, self.footprint / 2 * h / w])
        c = self.positions[i]
        lo, hi = c - half, c + half
        xs, ds = [], []
        for cx in range(int(np.floor(lo[0])), int(np.floor(hi[0])) + 1):
            for cy in range(int(np.floor(lo[1])), int(np.floor(hi[1])) + 1):
                xy, d = self._cell(cx, cy)
                keep = np.all((xy >= lo) & (xy < hi), axis=1)
                xs.append(xy[keep])
                ds.append(d[keep])
        xy = np.concatenate(xs)
        d = np.concatenate(ds)
        if len(xy) > self.n_desc:
            sel = rng.permutation(len(xy))[:self.n_desc]
            xy, d = xy[sel], d[sel]
        pix = (xy - c) * px_per_unit + np.array([w / 2, h / 2]) + rng.normal(0, self.pixel_noise, xy.shape)
        d = flip_bits(d, self.noise, rng)
        n_clutter = self.n_desc - len(d)
        if n_clutter > 0:
            pix = np.concatenate([pix, rng.random((n_clutter, 2)) * np.array([w, h])])
            d = np.concatenate([d, random_descriptors(n_clutter, rng)])
        strength = np.sort(rng.random(len(d)).astype(np.float32))[::-1].copy()
        order = rng.permutation(len(d))
        return d[order].copy(), pix[order].copy(), strength


def guided_visits(n_q, n_c, width=4000.0, height=3000.0, radius=150.0, match_fraction=0.6, noise=0.08, jitter=20.0,
                  seed=11):
    """Dense-stage guided matching workload (src/dense/dense_stereo.cpp:217-281) for ONE candidate image:
    n_c candidate features at uniform positions in a width x height image, n_q source features whose predicted
    position in that image is known; `match_fraction` of them are noisy copies of a candidate within `jitter`
    pixels of the prediction (so the 0.85 ratio rule passes for a controlled subset), the rest are unrelated.
    Queries are ordered along a Hilbert curve of the predicted position like the reference orders its source
    features (:23-48,191-193). Candidate lists = features within `radius` of the prediction by ascending distance
    (jk-tree's result order). -> dict(q, c, pred_xy, cand_xy, begin uint64 [n_q+1], nearby uint32)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    cand_xy = np.stack([rng.uniform(0, width, n_c), rng.uniform(0, height, n_c)], axis=1)
    c = random_descriptors(n_c, rng)
    q = random_descriptors(n_q, rng)
    m = int(n_q * match_fraction)
    src = rng.integers(0, max(n_c, 1), m) if n_c else np.zeros(0, np.int64)
    pred_xy = np.stack([rng.uniform(0, width, n_q), rng.uniform(0, height, n_q)], axis=1)
    if m and n_c:
        q[:m] = flip_bits(c[src], noise, rng)
        pred_xy[:m] = cand_xy[src] + rng.normal(0, jitter, (m, 2))
    # Hilbert order of the predictions (d2xy-style index on a 4096 grid)
    def hilbert(x, y, order=4096):
        x, y = x.astype(np.int64).copy(), y.astype(np.int64).copy()
        d = np.zeros(len(x), np.int64)
        s = order // 2
        while s > 0:
            rx, ry = ((x & s) > 0).astype(np.int64), ((y & s) > 0).astype(np.int64)
            d += s * s * ((3 * rx) ^ ry)
            flip = (ry == 0) & (rx == 1)
            x, y = np.where(flip, s - 1 - x, x), np.where(flip, s - 1 - y, y)
            swap = ry == 0
            x, y = np.where(swap, y, x), np.where(swap, x, y)
            s //= 2
        return d
    perm = np.argsort(hilbert(np.clip(pred_xy[:, 0], 0, width - 1), np.clip(pred_xy[:, 1], 0, height - 1)), kind="stable")
    q, pred_xy = q[perm], pred_xy[perm]
    begin = np.zeros(n_q + 1, np.uint64)
    nearby = np.zeros(0, np.uint32)
    if n_c and n_q:
        pairs = cKDTree(pred_xy).sparse_distance_matrix(cKDTree(cand_xy), radius * (1 + 1e-9), output_type="ndarray")
        owner, flat = pairs["i"].astype(np.int64), pairs["j"].astype(np.int64)
        d2 = ((cand_xy[flat] - pred_xy[owner]) ** 2).sum(axis=1)
        keep = d2 < radius * radius  # jk-tree keeps maxRadius > distance (squared), strictly
        flat, owner, d2 = flat[keep], owner[keep], d2[keep]
        # per list: ascending distance, the KD-tree's result order (two single-key sorts: by distance, then a stable
        # one by list -- much faster than a three-key lexsort; exact distance ties do not occur for random positions)
        order = np.argsort(d2)
        order = order[np.argsort(owner[order], kind="stable")]
        nearby = flat[order].astype(np.uint32)
        begin[1:] = np.cumsum(np.bincount(owner, minlength=n_q)).astype(np.uint64)
    return dict(q=q, c=c, pred_xy=pred_xy, cand_xy=cand_xy, begin=begin, nearby=nearby.astype(np.uint32))
